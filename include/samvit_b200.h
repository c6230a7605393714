/* samvit_b200 — C ABI of the B200-native SAM ViT image-encoder forward path.
 *
 * The reference has no C/FFI operator API for this path: its boundary is the Python module
 * `sam.modeling.ImageEncoderViT` (sam/modeling/image_encoder.py:17-120) constructed by
 * `sam/build_sam.py:60-73`.  This header is the C boundary a host binds INSTEAD of that module's
 * forward; each entry names the reference interface it replaces.  Plain pointers and sizes only,
 * no torch types.  Every function returns 0 on success; on failure a non-zero code, and
 * svb_last_error() returns a thread-local message.  All device work is enqueued on the given
 * CUDA stream; nothing synchronises unless stated.  There is NO CPU fallback.
 */
#ifndef SAMVIT_B200_H
#define SAMVIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svb_encoder svb_encoder_t;
typedef void* svb_stream_t; /* cudaStream_t */

enum { SVB_MODE_BF16 = 0,      /* bf16 operands on tcgen05 tensor cores, fp32 accumulate / residual / norms  */
       SVB_MODE_FP32 = 1 };    /* validation mode: fp32 operands, fp32 FMA accumulate (1e-4 rel-L2 bar)       */
enum { SVB_DTYPE_F32 = 0, SVB_DTYPE_BF16 = 1, SVB_DTYPE_F16 = 2 /* inputs only */ };

/* Constructor arguments of ImageEncoderViT (image_encoder.py:18-36) as fixed by _build_sam (build_sam.py:60-73). */
typedef struct svb_config {
    int32_t img_size;          /* 1024 */
    int32_t patch_size;        /* 16 */
    int32_t in_chans;          /* 3 */
    int32_t embed_dim;         /* 768 / 1024 / 1280 */
    int32_t depth;             /* 12 / 24 / 32 */
    int32_t num_heads;         /* 12 / 16 / 16 */
    int32_t mlp_dim;           /* int(embed_dim * mlp_ratio) */
    int32_t window_size;       /* 14 */
    int32_t num_global;        /* len(global_attn_indexes) */
    int32_t global_idx[16];    /* global_attn_indexes */
    int32_t fpn_dims[4];       /* SimpleFPN out_dims = {128,256,512,1024} (image_encoder.py:105) */
    float ln_eps;              /* 1e-6 */
    float gn_eps;              /* 1e-5 */
} svb_config_t;

const char* svb_last_error(void);
int svb_version(void);

/* ---- encoder object: replaces ImageEncoderViT.__init__ / load_state_dict / forward ---- */
int svb_encoder_create(const svb_config_t* cfg, svb_encoder_t** out);
void svb_encoder_destroy(svb_encoder_t* enc);

/* Replaces nn.Module.load_state_dict for one entry: `key` is the reference state_dict key
 * (e.g. "blocks.3.attn.qkv.weight"); `data` is a DEVICE pointer to `numel` contiguous
 * fp32 values in the reference layout.  The data is repacked (bf16 and fp32 copies, ConvTranspose /
 * Conv2d(k=2,s=2) weights re-laid as GEMM operands) into memory owned by the encoder.  Keys of the
 * never-executed `orig_neck` (image_encoder.py:88-104) are accepted and ignored. */
int svb_encoder_load_param(svb_encoder_t* enc, const char* key, const float* data, int64_t numel, svb_stream_t stream);
/* Number of forward-path parameters not yet loaded (0 = ready). */
int svb_encoder_missing_params(const svb_encoder_t* enc);

/* Workspace the forward needs for `chunk` images processed at once (activations; no allocation inside forward). */
size_t svb_encoder_workspace_bytes(const svb_encoder_t* enc, int chunk, int mode);

/* Replaces ImageEncoderViT.forward (image_encoder.py:107-120): x is a DEVICE fp32 tensor (B,3,S,S) NCHW;
 * res2..res5 are DEVICE output tensors NCHW of dtype `out_dtype` with shapes (B,128,S/4,S/4), (B,256,S/8,S/8),
 * (B,512,S/16,S/16), (B,1024,S/32,S/32).  The batch is processed in chunks of `chunk` images. */
int svb_encoder_forward(svb_encoder_t* enc, const float* x, int batch, void* res2, void* res3, void* res4, void* res5,
                        int out_dtype, int mode, int chunk, void* workspace, size_t workspace_bytes, svb_stream_t stream);

/* Scope row N3: the forward on canvases other than img_size x img_size (both sides multiples of 32 * patch_size, e.g. the
 * 1024 x 2048 pads of the reference's COCO evaluation).  Takes the reference's own fallbacks: `interpolate_pos_encoding`
 * (bicubic pos_embed, image_encoder.py:111-114,124-132) and, in the global-attention blocks, `get_rel_pos`'s linear resize of the
 * rel_pos tables (:319-330).  x (B,3,img_h,img_w); outputs (B,C_k,img_h/s_k,img_w/s_k).  The attention of such grids runs on the
 * fp32-math kernel (the tcgen05 attention kernels implement the trained 64 x 64 grid). */
size_t svb_encoder_workspace_bytes_hw(const svb_encoder_t* enc, int chunk, int mode, int img_h, int img_w);
int svb_encoder_forward_hw(svb_encoder_t* enc, const float* x, int batch, int img_h, int img_w, void* res2, void* res3, void* res4,
                           void* res5, int out_dtype, int mode, int chunk, void* workspace, size_t workspace_bytes, svb_stream_t stream);
/* The same forward for an input tensor of dtype `x_dtype` (SVB_DTYPE_F32 / _BF16 / _F16): the reference's pipeline hands the encoder
 * fp16 images after cast_batch_to_half (pipeline/XDecoderPipeline.py:93-95); the cast to the GEMM operand type happens in the
 * patch-embedding loader, no fp32 copy of the batch is made.  img_h x img_w as in svb_encoder_forward_hw (img_size x img_size takes
 * the same passes as svb_encoder_forward). */
int svb_encoder_forward_x(svb_encoder_t* enc, const void* x, int x_dtype, int batch, int img_h, int img_w, void* res2, void* res3,
                          void* res4, void* res5, int out_dtype, int mode, int chunk, void* workspace, size_t workspace_bytes,
                          svb_stream_t stream);
/* The two table fallbacks alone (device pointers, fp32): F.interpolate(pos_embed, mode='bicubic') (h0,w0,dim)->(h1,w1,dim);
 * F.interpolate(rel_pos, mode='linear') (len0,head_dim)->(len1,head_dim). */
int svb_resize_pos_embed(const float* src, float* dst, int h0, int w0, int h1, int w1, int dim, svb_stream_t stream);
int svb_resize_rel_pos(const float* src, float* dst, int len0, int len1, int head_dim, svb_stream_t stream);

/* The same forward fed with what the reference's callers hold BEFORE their eager pre-processing (scope row N2): per image a DEVICE
 * pointer to a uint8 (C,h,w) tensor, h,w <= img_size.  Replaces `(x - pixel_mean) / pixel_std` + `ImageList.from_tensors(images, 1024)`
 * (modeling/architectures/xdecoder_model.py:481-484, detectron2 zero padding to the bottom/right) + ImageEncoderViT.forward: the
 * normalisation and the padding happen inside the patch-embedding loader; no fp32 canvas is materialised.  `images`, `heights`,
 * `widths`, `pixel_mean`, `pixel_std` are HOST arrays (batch / batch / batch / in_chans / in_chans entries). */
int svb_encoder_forward_u8(svb_encoder_t* enc, const uint8_t* const* images, const int* heights, const int* widths,
                           const float* pixel_mean, const float* pixel_std, int batch, void* res2, void* res3, void* res4, void* res5,
                           int out_dtype, int mode, int chunk, void* workspace, size_t workspace_bytes, svb_stream_t stream);
/* The staging step alone: out = patch rows [batch * (img/patch)^2, chans * patch^2] of the normalised, zero-padded canvases. */
int svb_stage_images_u8(const uint8_t* const* images, const int* heights, const int* widths, int batch, int chans, int img, int patch,
                        const float* pixel_mean, const float* pixel_std, void* out, int out_dtype, svb_stream_t stream);

/* Same call with HOST buffers (the end-to-end path): copies each chunk's input host->device, runs the forward and
 * copies the four outputs device->host, double-buffered on internal streams; returns after everything completed.
 * Host buffers should be pinned for full PCIe bandwidth. */
int svb_encoder_forward_host(svb_encoder_t* enc, const float* x_host, int batch, void* res2_host, void* res3_host,
                             void* res4_host, void* res5_host, int out_dtype, int mode, int chunk);
/* The passes (images per pass through the kernels) svb_encoder_forward (host_path = 0) / svb_encoder_forward_host (host_path = 1) split
 * `batch` into for a given `chunk`: a dynamic programme over the wave quantisation of the block GEMMs; the host path also charges the
 * upload of the first pass and the download of the last one (the only copies its pipeline exposes).  Writes the sizes to passes[] and
 * returns 1000 + their number (error codes stay below 1000). */
int svb_encoder_pass_schedule(svb_encoder_t* e, int batch, int chunk, int host_path, int out_dtype, int* passes, int max_passes);
/* The same schedules from the geometry alone (no encoder handle, no device: host logic that the CPU tests exercise): embed_dim /
 * mlp_dim / depth of the blocks, tokens per image, SMs of the device, bytes per image the host path uploads / downloads. */
int svb_pass_schedule_model(int embed_dim, int mlp_dim, int depth, int tokens, int sms, int batch, int chunk, int host_path,
                            double in_bytes_per_image, double out_bytes_per_image, int* passes, int max_passes);

/* Token stream (B*T, D) fp32 after the patch embedding (block = -1) or after block `block` of the LAST forward's
 * last chunk, copied into `dst` (device).  Only valid when taps were enabled; used by the parity tests to bisect. */
int svb_encoder_enable_taps(svb_encoder_t* enc, int enable);
int svb_encoder_read_tap(svb_encoder_t* enc, int block, float* dst, int64_t numel, svb_stream_t stream);

/* ---- single operators (the pieces of the forward; exported so each can be checked against the oracle) ---- */
/* nn.Linear / conv-as-GEMM: C[M,N] = act(A[M,K] W[N,K]^T + bias) + resid.  mode BF16: A,W bf16 (tcgen05); FP32: fp32.
 * act_gelu: 0 none, 1 exact-erf GELU (nn.GELU()), 2 ReLU (F.relu: the FFN of the deformable encoder layer). */
int svb_linear(int mode, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act_gelu,
               const float* resid, int ldr, int resid_mod, void* out, int out_dtype, int ldo, double* gn_stats,
               int rows_per_sample, int remap_grid, int remap_grid_pad, svb_stream_t stream);
/* out_t[n * ldo + m] = sum_k A[m, k] W[n, k] + bias[n]: svb_linear (bf16 operands, fp32 accumulate) with the fp32 result stored
 * TRANSPOSED.  For products with few W rows and many A rows — the mask logits `torch.einsum("bqc,bchw->bqhw", mask_embed,
 * mask_features)` of XDecoder.forward_prediction_heads (modeling/interface/xdecoder.py:459): A = the image's mask-feature rows
 * (h*w, c), W = its 101 mask embeddings (q, c), out_t = outputs_mask[b] (q, h*w) — the long dimension runs along the GEMM's M, so
 * no tile rows are spent on padding the 101 queries to 256.  tcgen05 path only (M % 4 == 0, ldo % 8 == 0, 16-byte aligned operands). */
int svb_linear_nt(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, float* out_t, int ldo,
                  svb_stream_t stream);
/* The same GEMM with the LayerNorm-folding epilogues of the bf16 path: norm1 / norm2
 * (image_encoder.py:183,195) never run as separate passes.
 *   producer side (fp32 `out` + `resid` required): `out_bf16_copy` (bf16 [M, ldo2]) receives a rounded copy of the final rows and
 *     `stat_out` ([M][ceil(N/128)][2] fp32) the partial (sum, sum of squares) of each final row per 128-column slab;
 *   consumer side: `ln_stats` = such partials of the rows of x (over ln_dim = K columns), A = bf16(x) UN-normalised,
 *     W / bias / ln_colsum as written by svb_fold_layernorm:  out = act(rstd (A W^T - mu ln_colsum) + bias). */
int svb_linear_fused(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act_gelu,
                     const float* resid, int ldr, int resid_mod, void* out, int out_dtype, int ldo, const float* ln_stats,
                     const float* ln_colsum, int ln_dim, float ln_eps, void* out_bf16_copy, int ldo2, float* stat_out,
                     int remap_grid, int remap_grid_pad, svb_stream_t stream);
/* Derive the folded operands of `LayerNorm(gamma, beta) -> Linear(W [N,K], bias)` from fp32 masters (device pointers):
 * Wg bf16 [N,K] = gamma (.) W, colsum [N] = row sums of Wg, bias_f [N] = bias + W beta. */
int svb_fold_layernorm(const float* W, const float* bias, const float* gamma, const float* beta, void* Wg_bf16, float* colsum,
                       float* bias_f, int N, int K, svb_stream_t stream);
/* remap_grid > 0: output row r (token order, grid x grid per image) is stored at the token's row of the window-padded
 * grid_pad x grid_pad layout — window_partition's F.pad (image_encoder.py:271-275) expressed as a store address. */
/* out = cast(a [+ b]) elementwise (a, b fp32 device pointers, b may be NULL): `with_pos_embed` + the cast to the GEMM operand type
 * (transformer_encoder_deform.py:112-114,124). */
int svb_add_cast(const float* a, const float* b, void* out, int out_dtype, int64_t numel, svb_stream_t stream);
/* nn.LayerNorm over the last dim (image_encoder.py:166,176): fp32 in, out dtype = out_dtype.  With `add` (rows x dim, element
 * type = out_dtype) the residual add of Block.forward (image_encoder.py:194) is fused in: x += add is written back, then
 * out = LayerNorm(x). */
int svb_layernorm(float* x, const void* add, const float* weight, const float* bias, void* out, int out_dtype, int rows, int dim,
                  float eps, svb_stream_t stream);
/* Block attention core (image_encoder.py:239-255 + 258-304 + 340-376): qkv [B*g*g, 3*D] token order ->
 * out [B*g*g, D]; ws == g selects global attention, otherwise ws x ws windows with bias-valued pad keys.
 * impl: 0 = fp32-math SIMT kernel (dtype f32 or bf16), 1 = tcgen05 kernel (bf16 only). */
int svb_attention(int impl, int dtype, const void* qkv, void* out, const float* rel_pos_h, const float* rel_pos_w,
                  const float* qkv_bias, int batch, int grid, int ws, int heads, int head_dim, svb_stream_t stream);
/* tcgen05 attention core (bf16).  Same computation as svb_attention; the operands are in the layout the encoder keeps:
 *   qkv      ws == grid: token order [B*grid*grid, 3*D];  ws == 14: window-padded [B, 70, 70, 3*D] whose pad rows hold
 *            the qkv bias (svb_fill_pad_rows) — window_partition / window_unpartition (image_encoder.py:258-304) are
 *            TMA box coordinates and store addresses, nothing is copied;
 *   rel_pack bf16 [svb_rel_pack_rows(ws, grid)][head_dim] written by svb_pack_rel_table from rel_pos_h / rel_pos_w;
 *   out      token order [B*grid*grid, D] bf16.
 * Implemented geometry: grid 64, ws 14 or 64, head_dim 64 or 80 (everything build_sam.py:14-44 instantiates). */
int svb_attention_tc(const void* qkv, void* out, const void* rel_pack, int batch, int grid, int ws, int heads, int head_dim,
                     svb_stream_t stream);
/* Same call with per-phase cycle counters of the softmax warp groups (debug aid; `phase_clocks` = 16 device int64, or NULL). */
int svb_attention_tc_phases(const void* qkv, void* out, const void* rel_pack, int batch, int grid, int ws, int heads,
                            int head_dim, long long* phase_clocks, svb_stream_t stream);
int svb_rel_pack_rows(int ws, int grid);
/* Debug aid: a host-mapped (zero-copy) buffer of 64 uint64 that receives {block, thread, barrier address, parity} records when
 * an mbarrier wait of the attention kernels times out (the kernel then traps). */
int svb_attention_debug_buffer(void* mapped_device_ptr);
int svb_pack_rel_table(const float* table, void* rel_pack, int table_len, int head_dim, int is_w, svb_stream_t stream);
int svb_fill_pad_rows(void* qkv_padded, const float* qkv_bias, int batch, int grid, int grid_pad, int row_len,
                      svb_stream_t stream);
/* Scope row N3 on the tensor cores: the same attention cores for token grids other than 64 x 64 (both sides multiples of 32; e.g. the
 * 64 x 128 grid of the reference's 1024 x 2048 evaluation pads).  Windowed blocks: qkv on the window-padded grid [B, 14 ceil(gh / 14),
 * 14 ceil(gw / 14), 3D] (pad rows = qkv bias: svb_fill_pad_rows_hw), rel_pack as for svb_attention_tc (ws 14).  Global blocks: qkv in
 * token order; rel_pos_h / rel_pos_w are the fp32 tables ALREADY resized to 2 gh - 1 / 2 gw - 1 rows (get_rel_pos's linear resize,
 * image_encoder.py:319-330 = svb_resize_rel_pos); the decomposed terms q . rel_pos[j] are formed per head by the tcgen05 GEMM into the
 * workspace, then the generic-grid kernel (csrc/attention_ext.cu) runs. */
int svb_fill_pad_rows_hw(void* qkv_padded, const float* qkv_bias, int batch, int grid_h, int grid_w, int row_len, svb_stream_t stream);
int svb_attention_window_hw(const void* qkv_padded, void* out, const void* rel_pack, int batch, int grid_h, int grid_w, int heads,
                            int head_dim, svb_stream_t stream);
size_t svb_attention_global_hw_workspace(int batch, int grid_h, int grid_w, int heads, int head_dim);
int svb_attention_global_hw(const void* qkv, void* out, const float* rel_pos_h, const float* rel_pos_w, int batch, int grid_h, int grid_w,
                            int heads, int head_dim, void* workspace, size_t workspace_bytes, svb_stream_t stream);
/* PatchEmbed im2col (image_encoder.py:402-410). */
int svb_im2col(const float* x, void* out, int out_dtype, int batch, int chans, int img, int patch, svb_stream_t stream);
/* GroupNorm(1,C) apply from (sum,sumsq) statistics; NHWC rows -> NHWC rows, or -> NCHW with `levels` folded 2x2 stages. */
int svb_groupnorm_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* out, int out_dtype,
                        int64_t rows, int C, int64_t rows_per_sample, float eps, int gelu, svb_stream_t stream);
int svb_groupnorm_apply_nchw(const float* x, const double* stats, const float* gamma, const float* beta, void* out,
                             int out_dtype, int batch, int grid, int levels, int C, float eps, int gelu, svb_stream_t stream);

/* ---- scope row N1: the first consumer of res3..res5 (MSDeformAttn pixel decoder) ----
 * Replaces MSDA.ms_deform_attn_forward, the reference's only native kernel (modeling/vision/encoder/ops/src/cuda/
 * ms_deform_attn_cuda.cu / ms_deform_im2col_cuda.cuh:242-303), as MSDeformAttnFunction.forward calls it
 * (ops/functions/ms_deform_attn_func.py:34-40): value (N,S,M,D), sampling_locations (N,Lq,M,L,P,2) in [0,1] (x,y),
 * attention_weights (N,Lq,M,L,P) -> out (N,Lq,M*D).  `dtype` is the element type of value and out (fp32, or bf16 with fp32
 * accumulation — the reference kernel is fp32-only); locations and weights are fp32.  `spatial_shapes` ([L][2] = (H,W)) and
 * `level_start_index` ([L]) are HOST arrays; the levels must tile the S positions.  (The reference's im2col_step batching
 * argument has no effect on the result and is not needed.) */
int svb_ms_deform_attn_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                               const float* sampling_locations, const float* attention_weights, void* out, int dtype, int batch,
                               int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_points,
                               svb_stream_t stream);

/* The sampling core with the head of MSDeformAttn.forward (ops/modules/ms_deform_attn.py:97-112) fused in: instead of materialised
 * sampling_locations / attention_weights it takes `offsets_and_logits` = the fp32 output rows [N*Lq, M*L*P*3] of the module's two
 * query Linears (M*L*P*2 sampling offsets laid out (M,L,P,2), then M*L*P attention logits laid out (M,L*P)) and
 * `reference_points` (N,Lq,L,ref_dim), ref_dim 2 (points) or 4 (boxes); the softmax over L*P and the location arithmetic happen
 * in the kernel. */
int svb_ms_deform_attn_fused_forward(const void* value, const int32_t* spatial_shapes, const int32_t* level_start_index,
                                     const float* reference_points, int ref_dim, const float* offsets_and_logits, void* out, int dtype,
                                     int batch, int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                                     int num_points, svb_stream_t stream);

/* ---- measurement helpers (bench.py): launch accounting and a CUDA-event profiler.  Between start and stop every kernel
 * launch of this library is bracketed by events on its launch stream; stop synchronises the device and returns, per
 * category {0 GEMM, 1 windowed attention, 2 global attention, 3 norms, 4 other}, the summed device milliseconds,
 * algorithmic FLOPs, algorithmic bytes and launch counts. ---- */
int svb_profile_start(void);
int svb_profile_stop(double* ms5, double* flops5, double* bytes5, int64_t* launches5);
int64_t svb_launch_count(void);   /* kernels launched by this library since it was loaded */

/* ---- scope row N1, convolutional part of the pixel decoder (`MSDeformAttnPixelDecoder.forward`,
 * modeling/vision/encoder/transformer_encoder_deform.py:315-359).  "rows" = [sample][pixel][channel] (NHWC), the layout the GEMM
 * reads and writes: a 1x1 convolution (`input_proj[i][0]` :209-214, `lateral_conv` :256-258, `mask_features` :238-245) is one
 * svb_linear on rows, the 3x3 `output_conv` (:259-268) svb_im2col3x3_rows + svb_linear. ---- */
/* (batch, channels, pixels) NCHW of src_dtype -> rows of dst_dtype; dst_sample_stride = elements between samples (0: dense), so that
 * the levels of `src_flatten` (:71,79) can be written in place. */
int svb_nchw_to_rows(const void* src, int src_dtype, void* dst, int dst_dtype, int batch, int channels, int pixels,
                     int64_t dst_sample_stride, svb_stream_t stream);
/* (batch, channels, pixels) NCHW -> sequence-first (pixels, batch, channels) [+ chan_add per channel, may be NULL]: `x.flatten(2) +
 * level_embed[i][None, :, None]` -> `.permute(2, 0, 1)` of the X-Decoder (modeling/interface/xdecoder.py:205-209) in one pass. */
int svb_nchw_to_seq(const void* src, int src_dtype, void* dst, int dst_dtype, int batch, int channels, int pixels, const float* chan_add,
                    svb_stream_t stream);
/* fp32 rows -> fp32 NCHW (`.transpose(1, 2).view(bs, -1, h, w)` :335 and the module's NCHW outputs :359). */
int svb_rows_to_nchw(const float* src, int64_t src_sample_stride, float* dst, int batch, int channels, int pixels, svb_stream_t stream);
/* nn.GroupNorm(groups, channels) on fp32 rows (+ optional ReLU): `input_proj[i][1]` (:213), detectron2 `get_norm("GN", C)` =
 * GroupNorm(32, C) of the lateral / output convs (:253-268).  stats_ws: batch * groups * 2 doubles of scratch. */
int svb_groupnorm_rows(const float* x, int64_t x_sample_stride, const float* gamma, const float* beta, void* out, int out_dtype,
                       int64_t out_sample_stride, int batch, int pixels, int channels, int groups, float eps, int relu,
                       double* stats_ws, svb_stream_t stream);
/* dst (batch, out_h, out_w, channels) += F.interpolate(src (batch, h, w, channels), size=(out_h, out_w), mode="bilinear",
 * align_corners=False) (:348). */
int svb_upsample_add_rows(const float* src, int64_t src_sample_stride, float* dst, int batch, int h, int w, int out_h, int out_w,
                          int channels, svb_stream_t stream);
/* im2col of a 3x3 / stride 1 / pad 1 convolution on fp32 rows: dst[b, y, x, (ky, kx, c)] (dst_dtype) — the A operand of `output_conv`. */
int svb_im2col3x3_rows(const float* src, void* dst, int dst_dtype, int batch, int h, int w, int channels, svb_stream_t stream);
/* The 3x3 `output_conv` (:259-268) without an im2col operand (bf16 path): src fp32 rows (batch, h, w, cin) are cast into a zero-padded bf16
 * map `padded_ws` (batch * (h + 2) * (w + 2) * cin bf16 elements of scratch) and the tcgen05 GEMM reads its A tiles from it with one row
 * shift per filter tap (implicit GEMM); weight_bf16 [cout, 9 * cin] ordered (ky, kx, c) as for svb_im2col3x3_rows; out fp32 rows
 * (batch, h, w, cout) = [relu](conv + bias).  w must be a multiple of 128, cin of 64, cout of 32. */
int svb_conv3x3_rows(const float* src, const void* weight_bf16, const float* bias, float* out, void* padded_ws, int batch, int h, int w,
                     int cin, int cout, int relu, svb_stream_t stream);
/* Post-norm LayerNorm of the deformable encoder layer in one pass (transformer_encoder_deform.py:126-127 `src = norm1(src + src2)`,
 * :119 `src = norm2(src + ffn)`, followed by `with_pos_embed` :112-114 of the next layer): y = LayerNorm(x [+ add]) over rows of
 * `dim` (256 / 512 / 768 / 1024 / 1280); x and add (fp32, may be NULL) are only read.  Outputs, each optional (NULL): out = y (fp32),
 * out_bf16 = bf16(y), out_q_bf16 = bf16(y + pos[row mod pos_rows]) with pos fp32 (pos_rows x dim). */
int svb_layernorm_post(const float* x, const float* add, const float* weight, const float* bias, float* out, void* out_bf16,
                       const float* pos, int pos_rows, void* out_q_bf16, int rows, int dim, float eps, svb_stream_t stream);
/* One FPN level between its lateral convolution and the output of its 3x3 `output_conv` (transformer_encoder_deform.py:346-349), bf16
 * path: out (batch*h*w, cout) fp32 = act(conv3x3(GroupNorm(lateral) + F.interpolate(cur, (h, w), "bilinear")) + bias).  lateral
 * (batch*h*w, cin) fp32 = the raw lateral convolution; gn_gamma / gn_beta NULL: no norm; cur (batch, cur_h, cur_w, cin) fp32 rows with
 * `cur_sample_stride` elements between samples (<= 0: dense); padded_ws: batch*(h+2)*(w+2)*cin bf16; stats_ws: batch*groups*2 doubles.
 * The GroupNorm apply, the upsample-add and the zero-padded bf16 operand of the implicit GEMM are ONE pass (svb_groupnorm_rows +
 * svb_upsample_add_rows + svb_conv3x3_rows compute the same thing through two fp32 maps). */
int svb_fpn_conv3x3_rows(const float* lateral, const float* gn_gamma, const float* gn_beta, int gn_groups, float gn_eps, const float* cur,
                         int64_t cur_sample_stride, int cur_h, int cur_w, const void* weight_bf16, const float* bias, float* out,
                         void* padded_ws, double* stats_ws, int batch, int h, int w, int cin, int cout, int relu, svb_stream_t stream);
/* out = cast(a + b[i mod b_numel]): svb_add_cast with a `b` shared by every sample (the sine position embedding + level embedding,
 * :73-75). */
int svb_add_cast_bcast(const float* a, const float* b, int64_t b_numel, void* out, int out_dtype, int64_t numel, svb_stream_t stream);

/* ---- scope row N4, first slice: the mask branch of `XDecoder.forward_prediction_heads` (modeling/interface/xdecoder.py:429-470).
 * `mask_embed` (MLP, :458) and the mask logits `einsum("bqc,bchw->bqhw")` (:459) are svb_linear calls. ---- */
/* x (batch, queries, channels) fp32, in place: the class token (row queries-1) <- sum_j softmax_j(cos(x_cls, x_j)) x_j over the object
 * tokens j < queries-1 (:440-446). */
int svb_cls_token_recompute(float* x, int batch, int queries, int channels, svb_stream_t stream);
/* F.interpolate(src (maps, h, w), size=(out_h, out_w), mode="bicubic", align_corners=False, antialias=True) (:463); tmp: maps * h * out_w
 * floats of scratch. */
int svb_resize_bicubic_aa(const float* src, float* tmp, float* dst, int maps, int h, int w, int out_h, int out_w, svb_stream_t stream);
/* out (batch, heads, per_sample) bool = sigmoid(v (batch, per_sample)) < 0.5, repeated over the heads (:467). */
int svb_mask_threshold_heads(const float* v, void* out_bool, int batch, int heads, int64_t per_sample, svb_stream_t stream);

/* The normalisation inside LanguageEncoder.compute_similarity (modeling/language/vlpencoder.py:242,244), which forward_prediction_heads
 * applies to the class embeddings (xdecoder.py:453-455): out[r, :] = scale * x[r, :] / (|x[r, :]| + eps); x fp32, out of `out_dtype`.
 * (The class-embedding projection, the similarity against the text embeddings and the box MLP are svb_linear calls.) */
int svb_l2_normalize_rows(const float* x, void* out, int out_dtype, int rows, int dim, float eps, float scale, svb_stream_t stream);

/* svb_mask_threshold_heads with the first statement of the next decoder layer folded in (xdecoder.py:467 + :267): v (batch, queries,
 * keys) logits -> out (batch * heads, queries, keys) bool = sigmoid(v) < 0.5 repeated over the heads, except that a row whose EVERY key
 * would be masked is written all-False (`attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False`). */
int svb_mask_threshold_heads_clear(const float* v, void* out_bool, int batch, int heads, int queries, int keys, svb_stream_t stream);

/* ---- scope row N4, second slice: the attention core of `CrossAttentionLayer.forward_post` (modeling/interface/modules.py:95-106), i.e.
 * nn.MultiheadAttention's softmax((q / sqrt(d)) k^T + mask) v for `queries` <= 128 tokens over `keys` image positions.  q (queries, batch,
 * heads * 64), k / v (keys, batch, heads * 64) sequence-first as the reference passes them, element type `dtype`; mask_bool
 * (batch * heads, queries, keys) bytes, non-zero = not allowed (may be NULL); out (queries, batch, heads * 64) of `dtype`.  workspace:
 * svb_masked_cross_attention_workspace(...) floats.  A row whose keys are all masked yields NaN, as torch does. ---- */
int64_t svb_masked_cross_attention_workspace(int queries, int keys, int batch, int heads);
int svb_masked_cross_attention(const void* q, const void* k, const void* v, int dtype, const void* mask_bool, void* out, float* workspace,
                               int64_t workspace_floats, int queries, int keys, int batch, int heads, int head_dim, svb_stream_t stream);

/* mask (rows, keys) bool, in place: a row whose every entry is set is cleared (xdecoder.py:258). */
int svb_mask_clear_full_rows(void* mask_bool, int64_t rows, int keys, svb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SAMVIT_B200_H */
