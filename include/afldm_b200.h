/*
 * afldm_b200 - C ABI of the B200-native (sm_100a) AF-LDM denoising hot path.
 *
 * Every entry point takes plain device pointers + sizes + a CUDA stream, never allocates,
 * keeps no mutable global device state (graph-capturable, one process per GPU), and returns
 *      0            success
 *      < 0          AFLDM_E_* : argument/shape not supported (nothing was launched)
 *      > 0          a cudaError_t from the launch
 * Activations are fp32, channels-last ("NHWC": [B][H][W][C], C contiguous) unless a function
 * says NCHW.  This is the layout torch calls `channels_last`, so the reference's NCHW-shaped
 * tensors map onto it without a copy.
 *
 * Each function cites the reference interface it replaces (paths relative to /root/reference).
 * The reference attaches native code through torch-extension plugins
 * (afldm/af_libs/torch_utils/custom_ops.py:59-155, ops/upfirdn2d.cpp:16,102-105); the binding a
 * maintainer would add on the reference side is the ctypes stub shown in INTEGRATION.md.
 */
#ifndef AFLDM_B200_H
#define AFLDM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define AFLDM_API __attribute__((visibility("default")))
#else
#define AFLDM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* afldm_stream_t; /* cudaStream_t */

enum {
    AFLDM_OK = 0,
    AFLDM_E_SHAPE = -1,    /* size / divisibility not supported by the kernel family     */
    AFLDM_E_ARG = -2,      /* null pointer, bad enum, misaligned pointer                 */
    AFLDM_E_NOKERNEL = -3, /* no specialisation for this size (cf. filtered_lrelu.cpp:52) */
    AFLDM_E_WORKSPACE = -4 /* caller scratch too small                                    */
};

enum { AFLDM_ACT_IDENTITY = 0, AFLDM_ACT_SILU = 1 };
enum { AFLDM_CONV_SIMT_F32 = 0, AFLDM_CONV_TCGEN05_TF32 = 1, AFLDM_CONV_TCGEN05_F16 = 2 /* plan queries of afldm_conv2d_f16in_f32 */ };
enum { AFLDM_ATTN_SIMT_F32 = 0, AFLDM_ATTN_MMA_TF32 = 1 };

/* Library / build identification. abi = 1. */
AFLDM_API int afldm_abi_version(void);
/* Number of kernels this library has launched since load (host-side counter, for bench.py). */
AFLDM_API unsigned long long afldm_launch_count(void);
/* Human-readable text for a return code of this library (negative) or of CUDA (positive). */
AFLDM_API const char* afldm_error_string(int code);

/* ---- ideal resampling ----------------------------------------------------------------------
 * For a circular signal of even length n the reference's FFT filters are circulant:
 *   UpsampleRFFT(2):   y[2i] = x[i],  y[2i+1] = sum_j d[(i-j) mod n] x[j]
 *   LPF_RFFT(.5)[::2]: y[i]  = sum_m g[(2i-m) mod 2n] x[m]            (x has length 2n)
 * (afldm/af_libs/ideal_lpf.py:12-49 masks, :69-93, :112-134, :148-158).  The taps d, g are
 * baked into the library for n in {2,4,8,16,32,64,128} (afldm_b200/csrc/gen_taps.py).
 * Planes must be square (the reference builds its mask from the last dim only,
 * ideal_lpf.py:81-88) and C a multiple of 32.  Planes up to 32 x 32 run fused in one kernel; 64 and
 * 128 run as three line passes with fp32 intermediates in `workspace`
 * (afldm_resample_workspace_floats(op, ...) floats; op: 0 filtered act, 1 up2, 2 lpf_down2 with
 * H, W the sizes of the SMALL plane; 0 when no workspace is needed, NULL is then accepted). */
AFLDM_API size_t afldm_resample_workspace_floats(int op, int B, int H, int W, int C);

/* WarpedNonlinearity.forward (afldm/af_modules/af_blocks.py:19-28) for 4-D input:
 *   y = LPF(act(Up2(x * scale + shift)))[::2, ::2]   per (b, c) plane,
 * scale/shift are optional per-(b,c) vectors [B*C] (GroupNorm folded in: scale = gamma*rstd,
 * shift = beta - mean*gamma*rstd); pass NULL for none.  x, y: NHWC [B,H,W,C]; may alias. */
AFLDM_API int afldm_filtered_act_f32(const float* x, float* y, int B, int H, int W, int C, int act,
                           const float* scale, const float* shift, float* workspace,
                           size_t workspace_floats, afldm_stream_t stream);

/* afldm_filtered_act_f32 with an fp16 result (y: IEEE binary16 NHWC; no in-place form): see
 * afldm_filtered_act_gn_f16out below for why.  AFLDM_E_NOKERNEL where only an exact-FMA kernel applies
 * (AFLDM_FACT_MMA=0 at n >= 32). */
AFLDM_API int afldm_filtered_act_f16out(const float* x, void* y, int B, int H, int W, int C, int act,
                                        const float* scale, const float* shift, float* workspace,
                                        size_t workspace_floats, afldm_stream_t stream);

/* The tcgen05 form of the filtered activation (csrc/fact_tc.cu), selected explicitly: every 1-D circular convolution of
 * y = D act(U x U^T) D^T is a tcgen05.mma.kind::f16 GEMM with 128 lines as M (operands written by the threads into
 * SWIZZLE_128B K-major / MN-major tiles, accumulators in TMEM, 3-term fp16 split = fp32 accuracy).  Planes 16 x 16
 * (C % 16 == 0) and 32 x 32 (C % 8 == 0), 16-byte aligned x / y; AFLDM_E_NOKERNEL otherwise.  y_half != 0: y holds IEEE
 * binary16.  The default dispatch of afldm_filtered_act_* selects this kernel only under AFLDM_FACT_TC=1: on B200 the
 * warp-level mma.sync kernel is faster at these plane sizes (profiles/r02_fact_tc.md). */
AFLDM_API int afldm_filtered_act_tc(const float* x, void* y, int y_half, int B, int H, int W, int C, int act,
                                    const float* scale, const float* shift, afldm_stream_t stream);

/* The same with the GroupNorm finalised inside the kernel from the partial sums the producer of x emitted
 * (afldm_conv2d_f32 gn_partial; channels [0,Ca) from partial_a, [Ca,Ca+Cb) from partial_b as in
 * afldm_groupnorm_finalize_f32): GroupNorm -> filtered activation costs ONE launch and one pass over x.
 * Planes up to 32 x 32 (AFLDM_E_NOKERNEL above: finalize + afldm_filtered_act_f32). */
AFLDM_API int afldm_filtered_act_gn_f32(const float* x, float* y, int B, int H, int W, int C, int act,
                                        const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                        int slots_b, int Cb, int groups, float eps, const float* gamma,
                                        const float* beta, afldm_stream_t stream);
/* The same for the input torch.cat([xa, xb], dim=1) of an up-block resnet WITHOUT materialising the concat:
 * channels [0,Ca) are read from xa (NHWC, pitch Ca), [Ca,Ca+Cb) from xb (pitch Cb); y is the dense NHWC result
 * with Ca+Cb channels.  AFLDM_E_NOKERNEL when a CTA's channel group would straddle the two sources. */
AFLDM_API int afldm_filtered_act_gn_cat_f32(const float* xa, const float* xb, float* y, int B, int H, int W, int Ca,
                                  int Cb, int act, const float* partial_a, int slots_a,
                                  const float* partial_b, int slots_b, int groups, float eps,
                                  const float* gamma, const float* beta, afldm_stream_t stream);

/* afldm_filtered_act_gn_f32 / afldm_filtered_act_gn_cat_f32 with an fp16 result (y: IEEE binary16 NHWC, C halves per
 * pixel): the activation is consumed only by the resnet's next convolution (diffusers ResnetBlock2D conv1 / conv2,
 * SURVEY.md 8a-R), whose tensor-core products round their operands to 11 significant bits anyway - storing those 11
 * bits (fp16, round to nearest) halves the bytes the convolution stages (afldm_conv2d_f16in_f32).  |y| < 65504.
 * AFLDM_E_NOKERNEL where only the exact-FMA n = 32 kernel applies (AFLDM_FACT_MMA=0) and for planes above 32 x 32. */
AFLDM_API int afldm_filtered_act_gn_f16out(const float* x, void* y, int B, int H, int W, int C, int act,
                                           const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                           int slots_b, int Cb, int groups, float eps, const float* gamma,
                                           const float* beta, afldm_stream_t stream);
AFLDM_API int afldm_filtered_act_gn_cat_f16out(const float* xa, const float* xb, void* y, int B, int H, int W, int Ca,
                                               int Cb, int act, const float* partial_a, int slots_a,
                                               const float* partial_b, int slots_b, int groups, float eps,
                                               const float* gamma, const float* beta, afldm_stream_t stream);

/* UpsampleRFFT(up=2).forward (afldm/af_libs/ideal_lpf.py:148-158), optional affine on load:
 * x NHWC [B,H,W,C] -> y NHWC [B,2H,2W,C]. */
AFLDM_API int afldm_up2_ideal_f32(const float* x, float* y, int B, int H, int W, int C,
                        const float* scale, const float* shift, float* workspace,
                        size_t workspace_floats, afldm_stream_t stream);

/* afldm_up2_ideal_f32 with an fp16 result (y: IEEE binary16 NHWC [B,2H,2W,C]): the up-sampled tensor is consumed only
 * by the up-sampler's 3x3 convolution (af_blocks.py:99-104), see afldm_conv2d_f16in_f32.  workspace as in
 * afldm_up2_ideal_f32 (planes of 64 / 128; AFLDM_E_NOKERNEL there when the exact-FMA line passes are selected). */
AFLDM_API int afldm_up2_ideal_f16out(const float* x, void* y, int B, int H, int W, int C, float* workspace,
                                     size_t workspace_floats, afldm_stream_t stream);

/* LPF_RFFT(0.5)(x)[:, :, ::2, ::2] (afldm/af_modules/af_blocks.py:149-150):
 * x NHWC [B,2H,2W,C] -> y NHWC [B,H,W,C]  (H, W are the OUTPUT sizes). */
AFLDM_API int afldm_lpf_down2_f32(const float* x, float* y, int B, int H, int W, int C, float* workspace,
                        size_t workspace_floats, afldm_stream_t stream);
/* The same, also emitting the GroupNorm partial sums of y (gn_partial [B][1][C] float2 = (sum, sum of squares) per
 * output plane, the format afldm_groupnorm_finalize_f32 / afldm_filtered_act_gn_f32 consume with slots = 1), so the
 * resnet that follows a down-sampler needs no statistics pass.  Output planes up to 16 x 16. */
AFLDM_API int afldm_lpf_down2_gn_f32(const float* x, float* y, int B, int H, int W, int C, float* gn_partial,
                           afldm_stream_t stream);

/* ---- GroupNorm ----------------------------------------------------------------------------
 * torch.nn.GroupNorm(groups, C, eps) as used by diffusers ResnetBlock2D / Attention
 * (SURVEY.md 8a-R).  Produces the folded per-(b,c) affine  y = x*scale + shift :
 *   scale[b,c] = gamma[c]*rstd[b,g],  shift[b,c] = beta[c] - mean[b,g]*gamma[c]*rstd[b,g].
 * x NHWC [B,HW,C].  `partial` is caller scratch of afldm_groupnorm_scratch_floats(B,HW,C) floats
 * (0 in this build: one CTA per (b, group) reduces in a single launch; NULL is accepted). */
AFLDM_API size_t afldm_groupnorm_scratch_floats(int B, int HW, int C);
AFLDM_API int afldm_groupnorm_affine_f32(const float* x, int B, int HW, int C, int groups, float eps,
                               const float* gamma, const float* beta,
                               float* scale, float* shift, float* partial, afldm_stream_t stream);

/* The same scale / shift from partial sums emitted by the producer of x instead of a pass over x:
 * afldm_conv2d_f32 (tensor-core path) writes gn_partial [B][slots][C] float2 (sum, sum of squares over the
 * slot's pixels).  Channels [0,Ca) are taken from partial_a, [Ca,Ca+Cb) from partial_b - the two inputs
 * of a skip-connection torch.cat - pass Cb = 0 / NULL for a single source. */
AFLDM_API int afldm_groupnorm_finalize_f32(const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                           int slots_b, int Cb, int B, int HW, int groups, float eps,
                                           const float* gamma, const float* beta, float* scale, float* shift,
                                           afldm_stream_t stream);

/* y = act(x*scale[b,c] + shift[b,c]); x, y NHWC [B,HW,C] (plain nn.SiLU after a norm: UNet tail
 * conv_norm_out/conv_act, VAE blocks with up_filtered_act=false; act = identity gives the
 * normalised input of an attention block). scale/shift may be NULL. x may alias y. */
AFLDM_API int afldm_affine_act_f32(const float* x, float* y, int B, int HW, int C, int act,
                         const float* scale, const float* shift, afldm_stream_t stream);

/* afldm_affine_act_f32 with an fp16 result (y: IEEE binary16 [B,HW,C]; no in-place form): plain SiLU(GroupNorm(x))
 * in front of a tensor-core convolution (VAE blocks without the filtered activation). */
AFLDM_API int afldm_affine_act_f16out(const float* x, void* y, int B, int HW, int C, int act,
                                      const float* scale, const float* shift, afldm_stream_t stream);

/* GroupNorm (finalised from the producer's partial sums, as in afldm_groupnorm_finalize_f32) and
 * y = act(x*scale + shift) in one launch: the normalised input of an attention block
 * (diffusers Attention.group_norm in AttnProcessor2_0) without a separate finalize kernel. */
AFLDM_API int afldm_affine_act_gn_f32(const float* x, float* y, int B, int HW, int C, int act,
                            const float* partial_a, int slots_a, int Ca, const float* partial_b,
                            int slots_b, int Cb, int groups, float eps, const float* gamma,
                            const float* beta, afldm_stream_t stream);

/* afldm_affine_act_gn_f32 with an fp16 result (y: IEEE binary16 [B,HW,C]): the normalised attention input is consumed
 * only by the q | k | v projection (afldm_conv2d_f16in_f16out). */
AFLDM_API int afldm_affine_act_gn_f16out(const float* x, void* y, int B, int HW, int C, int act,
                                         const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                         int slots_b, int Cb, int groups, float eps, const float* gamma,
                                         const float* beta, afldm_stream_t stream);

/* ---- convolution / linear as implicit GEMM -----------------------------------------------
 * nn.Conv2d(Cin, Cout, k, stride=1, padding=k/2) with k in {1,3} on NHWC input, fused epilogue:
 *   y[b,h,w,:] = conv(x)[b,h,w,:] + bias + row_add[b,:] + residual[b,h,w,:]
 * (diffusers ResnetBlock2D conv1/conv2/conv_shortcut, Up/Downsample2D.conv with the stride
 *  forced to 1 by af_blocks.py:129, Attention to_q/to_k/to_v/to_out as k = 1).
 * x has row pitch x_pitch floats per pixel (>= Cin); w is PACKED [Cout][k*k][Cin] (tap-major,
 * Cin contiguous); bias [Cout] | NULL, row_add [B][Cout] with row pitch row_add_pitch | NULL
 * (time-embedding projection; a slice of the batched projection of all resnets),
 * residual NHWC with row pitch res_pitch | NULL (may alias y).  y has row pitch y_pitch
 * (>= Cout) so an output can be written straight into a channel slice of a concat buffer.
 * algo: AFLDM_CONV_SIMT_F32 = fp32 FMA (exact-fp32 class; any Cin/Cout),
 *       AFLDM_CONV_TCGEN05_TF32 = tcgen05 tensor cores, TF32 operands, fp32 accumulate in TMEM
 *       (same numeric class as the reference's default cuDNN path; falls back to
 *       AFLDM_E_NOKERNEL when the shape does not fit: Cin % 32, Cout % 16, W power of two).
 * workspace: split-K partial sums, afldm_conv2d_workspace_floats(...) floats (may be 0).
 * gn_partial: NULL, or [B][slots][Cout] float2 that receives the GroupNorm partial sums of y (the next
 *       layer's norm then needs no pass over y); slots = afldm_conv2d_gn_slots(...), 0 = not available
 *       for this shape / algo (then gn_partial must be NULL). */
AFLDM_API int afldm_conv2d_gn_slots(int B, int H, int W, int Cin, int Cout, int ksize, int algo);
/* 1 when `algo` has a kernel for this shape (host-side plan only, no launch): a producer asks this before it decides to
 * store an activation as fp16 for afldm_conv2d_f16in_f32 (algo = AFLDM_CONV_TCGEN05_F16). */
AFLDM_API int afldm_conv2d_supported(int B, int H, int W, int Cin, int Cout, int ksize, int algo);

/* Host-side query (no launch, works without a GPU): the launch plan of the tcgen05 convolution for this layer,
 * plan8 = {TMEM columns per CTA, dynamic shared-memory bytes per CTA, CTAs, CTA-pair mode (cta_group::2), tile width BN,
 * K splits, ring stages, halo mode}.  Exposes the invariant the library relies on: every plan asks for at least
 * 454 B of shared memory per TMEM column (1 KB per-CTA reserve included), so CTAs that share an SM (228 KB) never hold
 * more than 512 columns between them and tcgen05.alloc never blocks.  AFLDM_E_NOKERNEL when the
 * shape is outside the tcgen05 family. */
AFLDM_API int afldm_conv2d_plan(int B, int H, int W, int Cin, int Cout, int ksize, int algo, int* plan8);
AFLDM_API size_t afldm_conv2d_workspace_floats(int B, int H, int W, int Cin, int Cout, int ksize, int algo);
AFLDM_API int afldm_conv2d_f32(const float* x, int x_pitch, const float* w, const float* bias,
                     const float* row_add, int row_add_pitch, const float* residual, int res_pitch,
                     float* y, int y_pitch, int B, int H, int W, int Cin, int Cout, int ksize,
                     int algo, float* workspace, size_t workspace_floats, float* gn_partial,
                     afldm_stream_t stream);

/* afldm_conv2d_f32 (tensor-core path) on the input torch.cat([xa, xb], dim=1) without materialising it: input
 * channels [0,Ca) come from xa, [Ca,Ca+Cb) from xb (two TMA tensor maps, chosen per 32-channel chunk); w is packed
 * over Ca+Cb input channels.  The conv_shortcut of an up-block resnet.  AFLDM_E_NOKERNEL when Ca % 32 != 0 or the
 * shape is outside the tcgen05 family (callers then concatenate and use afldm_conv2d_f32). */
AFLDM_API int afldm_conv2d_cat_f32(const float* xa, int xa_pitch, int Ca, const float* xb, int xb_pitch, int Cb,
                         const float* w, const float* bias, const float* row_add, int row_add_pitch,
                         const float* residual, int res_pitch, float* y, int y_pitch, int B, int H, int W,
                         int Cout, int ksize, float* workspace, size_t workspace_floats, float* gn_partial,
                         afldm_stream_t stream);

/* The tensor-core convolution with an fp16 output (y: IEEE binary16, row pitch y_pitch halves, 8-byte aligned rows):
 * y = fp16(conv(x) + bias).  Used for the fused to_q | to_k | to_v projection in front of afldm_attention_f16.
 * AFLDM_E_NOKERNEL when the shape runs split-K or outside the tcgen05 family (callers then use afldm_conv2d_f32
 * and the fp32-input attention). */
AFLDM_API int afldm_conv2d_f16out(const float* x, int x_pitch, const float* w, const float* bias, void* y, int y_pitch,
                        int B, int H, int W, int Cin, int Cout, int ksize, afldm_stream_t stream);

/* afldm_conv2d_f32 (tensor-core path) with fp16 OPERANDS: x is IEEE binary16 NHWC (row pitch x_pitch halves, a multiple
 * of 8), w is the packed weight [Cout][k*k][Cin] rounded to binary16; tcgen05.mma.kind::f16, fp32 accumulation in TMEM,
 * fp32 epilogue and output exactly as afldm_conv2d_f32.  Same numeric class as AFLDM_CONV_TCGEN05_TF32 (products of
 * 11-bit significands, fp32 sums) at half the operand bytes and twice the K per instruction.  Cin % 64 == 0.
 * Plan queries: afldm_conv2d_workspace_floats / afldm_conv2d_gn_slots with algo = AFLDM_CONV_TCGEN05_F16.
 * AFLDM_E_NOKERNEL outside the family (callers keep fp32 activations and use afldm_conv2d_f32). */
AFLDM_API int afldm_conv2d_f16in_f32(const void* x, int x_pitch, const void* w, const float* bias,
                                     const float* row_add, int row_add_pitch, const float* residual, int res_pitch,
                                     float* y, int y_pitch, int B, int H, int W, int Cin, int Cout, int ksize,
                                     float* workspace, size_t workspace_floats, float* gn_partial,
                                     afldm_stream_t stream);

/* afldm_conv2d_f16out with fp16 operands as in afldm_conv2d_f16in_f32: y = fp16(conv(x) + bias), x, w, y binary16.
 * The fused to_q | to_k | to_v projection between afldm_affine_act_gn_f16out and afldm_attention_f16. */
AFLDM_API int afldm_conv2d_f16in_f16out(const void* x, int x_pitch, const void* w, const float* bias, void* y, int y_pitch,
                                        int B, int H, int W, int Cin, int Cout, int ksize, afldm_stream_t stream);

/* Backward of the filtered activation, elementwise part (the reference differentiates through
 * torch.fft in afldm/trainers/ldm_trainer.py:240-272): out[i] = g[i] * act'(z[i]) on the up-sampled plane, between the two
 * linear halves U^T (.) U and D^T (.) D, which run as afldm_plane_sep_transform_f32 with the transposed operators. */
AFLDM_API int afldm_act_bwd_mul_f32(const float* z, const float* g, float* out, long long n, int act, afldm_stream_t stream);

/* diffusers GEGLU, the feed-forward gate of BasicTransformerBlock in the SD-1.5 UNet2DConditionModel that
 * afldm/pipelines/video_equiv_editing_pipeline.py:680-686 evaluates: proj [rows][2H] -> y [rows][H],
 * y = proj[:, :H] * gelu(proj[:, H:]) with the exact (erf) GELU.  H % 4 == 0, 16-byte aligned pointers. */
AFLDM_API int afldm_geglu_f32(const float* proj, float* y, long long rows, int H, afldm_stream_t stream);

/* ---- cross-frame attention map store (afldm/pipelines/cross_frame_attn.py:78-97) -------------
 * CrossFrameAttnProcessor keeps, per attention layer, the layer input of the reference frame for every timestep
 * (`self.maps[store_id][t] = hidden_states`, keyed by the HOST value t.item() :31-33) and feeds it back as the K / V source
 * of later frames.  Here the maps of a layer are one device table [slots][n] and the slot (= step index) is read from
 * device memory, so STORE and LOAD passes can live in a captured CUDA graph:
 *   store != 0: table[*slot][0..n) = buf[0..n);   store == 0: buf[0..n) = table[*slot][0..n).
 * n % 4 == 0, 16-byte aligned pointers; the caller guarantees 0 <= *slot < slots. */
AFLDM_API int afldm_slot_copy_f32(float* table, float* buf, long long n, const int* slot, int store, afldm_stream_t stream);

/* ---- fractional shift (the equivariance measurement's warp) ---------------------------------
 * ImageShifter('ideal' | 'ideal_crop', r).shift(img, ti, tj) (afldm/shift_utils/shifters.py:157-191):
 * UpsampleRFFT(r) -> torch.roll by (round(ti r), round(tj r)) -> validity mask (gen_valid_mask :31-49) ->
 * [::r, ::r].  Every step is linear and separable, so per axis the chain is one matrix and the whole warp is
 *   y[p] = My[m] . x[p] . Mx[m]^T            per NCHW plane p, m = p / planes_per_matrix
 * with My [n_mat][Hout][Hin], Mx [n_mat][Wout][Win] built by the caller (fp64 -> fp32; afldm_b200/shift_utils).
 * A sweep of shifts of one image is ONE call (planes = shifts x channels), and the r-fold up-sampled tensor is never
 * materialised.  workspace: afldm_plane_sep_transform_workspace_floats(planes, Hin, Wout) floats. */
AFLDM_API size_t afldm_plane_sep_transform_workspace_floats(int planes, int Hin, int Wout);
AFLDM_API int afldm_plane_sep_transform_f32(const float* x, const float* my, const float* mx, float* y,
                                            float* workspace, size_t workspace_floats, int planes,
                                            int planes_per_matrix, int Hin, int Win, int Hout, int Wout,
                                            afldm_stream_t stream);

/* nn.Linear on a few rows (time embedding MLP, time_emb_proj): y[M,N] = act_in(x[M,K]) w[N,K]^T + b.
 * act_in applies SiLU to x on load (ResnetBlock2D: time_emb_proj(nonlinearity(temb))). M <= 64. */
AFLDM_API int afldm_linear_rows_f32(const float* x, const float* w, const float* bias, float* y,
                          int M, int K, int N, int act_in, int act_out, afldm_stream_t stream);

/* ---- attention ----------------------------------------------------------------------------
 * F.scaled_dot_product_attention(q,k,v) of diffusers AttnProcessor2_0 with the
 * cross-frame K/V source of afldm/pipelines/cross_frame_attn.py:79-97:
 *   q [B][Nq][heads*d] (row pitch q_pitch), k/v [Bkv][Nk][heads*d] (row pitch kv_pitch),
 *   batch b reads K/V batch  b / (B / Bkv)  (repeat semantics of :91-97), scale = d^-0.5,
 *   o [B][Nq][heads*d] (row pitch o_pitch).  d % 4 == 0, d <= 64 in this build.
 * algo: AFLDM_ATTN_SIMT_F32 = exact fp32 FMA (the reference's fp32 SDPA class);
 *       AFLDM_ATTN_MMA_TF32 = tensor-core products on TF32 operands (d % 8 == 0), fp32 accumulation
 *       and fp32 online softmax - used together with the TF32 convolution path. */
AFLDM_API int afldm_attention_f32(const float* q, int q_pitch, const float* k, const float* v, int kv_pitch,
                        float* o, int o_pitch, int B, int Bkv, int Nq, int Nk, int heads, int d,
                        int algo, afldm_stream_t stream);

/* The same attention with fp16 q / k / v (raw IEEE binary16, pitches in halves, rows 16-byte aligned, d % 8 == 0),
 * as written by afldm_conv2d_f16out; products on tensor cores with fp32 accumulation, fp32 softmax, fp32 output.
 * fp16 has the 11 significant bits of the TF32 operands of AFLDM_ATTN_MMA_TF32 (same numeric class) and a narrower
 * exponent: |q|, |k|, |v| must stay below 65504 (they are linear projections of GroupNorm outputs). */
AFLDM_API int afldm_attention_f16(const void* q, int q_pitch, const void* k, const void* v, int kv_pitch,
                        float* o, int o_pitch, int B, int Bkv, int Nq, int Nk, int heads, int d,
                        afldm_stream_t stream);
/* The same with an fp16 result (o: IEEE binary16, row pitch o_pitch halves): the attention output is consumed only by
 * the to_out projection (afldm_conv2d_f16in_f32). */
AFLDM_API int afldm_attention_f16_f16out(const void* q, int q_pitch, const void* k, const void* v, int kv_pitch,
                               void* o, int o_pitch, int B, int Bkv, int Nq, int Nk, int heads, int d,
                               afldm_stream_t stream);

/* In-place row softmax: x[r][:] = softmax(scale * x[r][:]) over `cols` entries, row pitch `pitch`.
 * Used by the large-head-dim attention (VAE mid block: 1 head of 512) which runs as
 * GEMM (q k^T) -> softmax -> GEMM (p v) through afldm_conv2d_f32 with ksize = 1. */
AFLDM_API int afldm_softmax_rows_f32(float* x, long long rows, int cols, int pitch, float scale,
                                     afldm_stream_t stream);

/* ---- small ops of one denoising step --------------------------------------------------------
 * diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): out[b] = [cos(t f_k) | sin(t f_k)],
 * f_k = exp(-ln(10000) k / (dim/2)). */
AFLDM_API int afldm_timestep_embedding_f32(const float* t, float* out, int B, int dim, afldm_stream_t stream);
/* torch.cat([a, b], dim=1) on NHWC: y[.., :Ca] = a, y[.., Ca:] = b (UNet up-path skip concat). */
AFLDM_API int afldm_concat_channels_f32(const float* a, int Ca, const float* b, int Cb, float* y,
                              long long pixels, afldm_stream_t stream);
/* y[.., :C] = x, y[.., C:Cpad] = 0 on NHWC (pixels = B*H*W): pads the 4-channel latents to the 32-channel box
 * granularity of the tensor-core convolution, so conv_in runs on tcgen05 (with zero-padded packed weights). */
AFLDM_API int afldm_pad_channels_f32(const float* x, int C, float* y, int Cpad, long long pixels, afldm_stream_t stream);
/* NCHW [B,C,H,W] <-> NHWC [B,H,W,C] (pipeline boundary: latents / decoded frames). */
AFLDM_API int afldm_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int HW, afldm_stream_t stream);
AFLDM_API int afldm_nhwc_to_nchw_f32(const float* x, float* y, int B, int C, int HW, afldm_stream_t stream);
/* DDIMScheduler.step with eta = 0 (afldm/pipelines/ldm_pipeline.py:103-109):
 *   x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t);  out = sqrt(a_p) x0 + sqrt(1-a_p) eps
 * passed as the two host-computed coefficients  out = cx*x + ce*eps.  out may alias x. */
AFLDM_API int afldm_axpby_f32(const float* x, const float* eps, float* out, float cx, float ce, long long n,
                    afldm_stream_t stream);

/* Same update with the two coefficients read from DEVICE memory (coef[0] = cx, coef[1] = ce), so a
 * captured CUDA graph of one denoising step can be replayed for every timestep. */
AFLDM_API int afldm_axpby_dev_f32(const float* x, const float* eps, float* out, const float* coef,
                                  long long n, afldm_stream_t stream);

/* ---- StyleGAN3 upfirdn2d (afldm/af_libs/torch_utils/ops/upfirdn2d.cpp:16; BASELINE config #1)
 * x NCHW [B,C,H,W] contiguous, f [fh][fw] (2-D, already normalised; gain applied here),
 * y NCHW [B,C,outH,outW], outW = (W*upx + padx0 + padx1 - fw + downx)/downx (upfirdn2d.cpp:35). */
AFLDM_API int afldm_upfirdn2d_f32(const float* x, const float* f, float* y, int B, int C, int H, int W,
                        int fh, int fw, int upx, int upy, int downx, int downy,
                        int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                        afldm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AFLDM_B200_H */
