/* dcl_b200.h -- C ABI of the B200-native sliding-window ClsWiseFormer inference path.
 *
 * The reference (mathwrx/Decouple-and-Couple_Learning_in_Multi-Modal_Brain_Tumor_Segmentation)
 * has no FFI: its boundary is two Python call signatures and a checkpoint schema
 * (SURVEY.md 8b).  Every entry point below names the reference call it replaces.  All
 * functions return 0 on success and a negative dcl_status otherwise; the message of the last
 * failure on the calling thread is returned by dcl_last_error().  No entry point falls back
 * to a CPU implementation: without a CUDA device every compute call fails with
 * DCL_ERR_CUDA.
 *
 * Conventions
 *   - tensors are fp32, NCDHW with the reference's axis order (N, C, X=240, Y=240, Z=155);
 *     Z is the contiguous axis (predict_overlap.py:34-41 slices it last)
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - a handle is bound to the CUDA device that was current in dcl_create and is not
 *     thread safe; use one handle per stream (different handles may be used from different threads)
 *   - environment: DCL_LANES=1..4 patches in flight per volume call (default 3), DCL_GATHER=0 accumulate-form stitch
 *     instead of the gather form, DCL_S2GEN=0 im2col GEMM instead of the rolling kernel for the stride-2 convolutions of
 *     the 64^3 level, DCL_PDL=1 programmatic dependent launch, DCL_STAMPS=1 / DCL_DEBUG=1 debug aids.
 *     Set CUDA_DEVICE_MAX_CONNECTIONS=32 before the CUDA context is created (the Python package does): with the
 *     driver's default of 8 hardware work queues the lane streams of a volume call can share a queue and serialise
 *     (measured: one process in four ran 33-42 instead of 25.5 ms per volume)
 *   - pointers named *_dev are device pointers, *_host host pointers; the caller owns all of them
 */
#ifndef DCL_B200_H
#define DCL_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define DCL_API __attribute__((visibility("default")))
#else
#define DCL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DCL_ABI_VERSION 1
#define DCL_PATCH 128            /* crop_H/W/D of test_overlap.py:48-52, img_dim of cls_wise_former.py:759 */
#define DCL_NUM_CLASSES 4        /* cls_wise_former.py:760 */
#define DCL_NUM_MODALITIES 4     /* cls_wise_former.py:762 */
#define DCL_KEEP_CHANNELS 16     /* InitConv out_channels, Unet_skipconnection.py:23 */
#define DCL_NUM_AUX 12           /* 4 heads x 3 regions, cls_wise_former.py:226-230 */

typedef enum dcl_status {
  DCL_OK = 0,
  DCL_ERR_ARG = -1,        /* bad argument (shape, NULL pointer, unknown name) */
  DCL_ERR_CUDA = -2,       /* CUDA runtime / launch failure, or no device */
  DCL_ERR_WEIGHTS = -3,    /* forward called before every required tensor was set */
  DCL_ERR_STATE = -4       /* handle used on the wrong device / after destroy */
} dcl_status;

/* Arithmetic mode of the convolution / GEMM kernels. */
typedef enum dcl_precision {
  DCL_FP32 = 0,            /* fp32 FFMA kernels: the parity mode (1e-3 rel. to the reference) */
  DCL_F16X3 = 1,           /* split operands on tcgen05: every activation / weight is a 16-bit hi + lo pair (fp16 + fp16 = 22
                              significant bits), every product a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (+ a_lo*w_lo) in the fp32
                              TMEM accumulator: the fast parity mode (same 1e-3 / 1e-4-label gates as DCL_FP32) */
  DCL_BF16X3 = 1,          /* the name the mode was specified under; same value */
  DCL_BF16 = 2             /* tcgen05 bf16 MMA, fp32 accumulate (2e-2 rel.) */
} dcl_precision;

/* Stitch / blend rule applied to the per-patch class probabilities. */
typedef enum dcl_stitch_mode {
  DCL_STITCH_REFERENCE = 0, /* predict_overlap.py:49-56 crop-and-overwrite, bug compatible (z tail shifted by 5) */
  DCL_STITCH_ALIGNED = 1,   /* same 8 corners, z tail taken from the aligned slice 101:128 */
  DCL_STITCH_UNIFORM = 2,   /* extension: sum(w*p)/sum(w), w = 1 (any patch list) */
  DCL_STITCH_GAUSSIAN = 3   /* extension: separable gaussian w, sigma = patch/8 */
} dcl_stitch_mode;

typedef struct dcl_config {
  int32_t abi_version;      /* DCL_ABI_VERSION */
  int32_t precision;        /* dcl_precision */
  int32_t want_aux;         /* 1: also run the 4 auxiliary heads (SuperviseLabel.py, EdgeSuperviseLabel.py) */
  int32_t keep_stages;      /* 1: keep named intermediate tensors readable through dcl_read_stage */
  int32_t reserved[12];
} dcl_config;

typedef struct dcl_handle dcl_handle;

DCL_API const char* dcl_last_error(void);
DCL_API int dcl_abi_version(void);

/* Replaces ClsWiseFormer.__init__ / get_cls_wise_former (cls_wise_former.py:43-278, :757-780):
 * allocates device workspace for one 128^3 patch.  fix_index.txt (:275-278) is not needed. */
DCL_API int dcl_create(const dcl_config* cfg, dcl_handle** out);
DCL_API int dcl_destroy(dcl_handle* h);
/* Bytes of device memory a handle allocates for `cfg` (dcl_workspace_bytes). */
DCL_API int64_t dcl_workspace_bytes(const dcl_config* cfg);

/* Replaces model.load_state_dict(checkpoint['state_dict']) (test_overlap.py:85-86).  `name` is
 * the reference state_dict key (an optional leading "module." is ignored); `data` is fp32,
 * host or device memory (detected), `numel` elements in the reference's layout.  Unknown names
 * and wrong sizes are errors.  Weights are repacked into kernel layouts once. */
DCL_API int dcl_set_weight(dcl_handle* h, const char* name, const float* data, int64_t numel);
/* The weight catalogue (pure host code, needs no device): the reference state_dict keys this library
 * accepts, their element counts, and whether they belong only to the auxiliary heads. */
DCL_API int dcl_weight_count(void);
DCL_API int dcl_weight_spec(int index, char* name, int32_t cap, int64_t* numel, int32_t* aux_only);
/* Number of state_dict tensors still missing for the configured path (0 = ready). */
DCL_API int dcl_missing_weights(dcl_handle* h, char* first_missing, int32_t cap);

/* Replaces ClsWiseFormer.forward(x, missing_modal) (cls_wise_former.py:585-592) for N = 1.
 *   x_dev        : first element of a (4,128,128,128) view; x_strides = element strides of
 *                  (C, X, Y, Z); the Z stride must be 1 (views of a volume are accepted as is,
 *                  predict_overlap.py:34-41)
 *   keep_scale_host : 16 floats, the dropout3d channel scale (0 or 1/0.8) of
 *                  Unet_skipconnection.py:31; NULL = deterministic (all ones)
 *   probs_dev    : (4,128,128,128) dense softmax probabilities (forward()[0])
 *   aux_dev      : NULL, or 12 dense (2,128,128,128) outputs ordered
 *                  supervise{01,02,04}, edge{01,02,04}, mid_semantic{..}, mid_edge{..}
 *                  (forward()[1..4]); requires cfg.want_aux */
DCL_API int dcl_forward(dcl_handle* h, const float* x_dev, const int64_t x_strides[4],
                const float* keep_scale_host, float* probs_dev, float* const* aux_dev, void* stream);

/* Replaces tailor_and_concat (predict_overlap.py:31-58) plus the label tail of validate_softmax
 * (:141-153) and utils/tools.py:89-109.
 *   vol_dev      : (4, X, Y, Z) dense fp32 volume, Z >= 155 for the reference modes
 *   starts_host  : n_patches x 3 patch origins (x,y,z); ignored (may be NULL) for
 *                  REFERENCE/ALIGNED, which use the 8 fixed corners
 *   keep_scale_host : n_patches x 16 or NULL
 *   probs_out_dev: NULL or (4, X, Y, Zout) stitched probabilities (Zout = 155 in reference
 *                  modes = y[..., :155], else Z)
 *   labels_out_dev: NULL or (X, Y, Zout) uint8 arg-max labels
 *   target_dev   : NULL or (X, Y, Zout) uint8 ground truth with label 4 already mapped to 3
 *   counts_out_dev: NULL or 13 uint64: label histogram [0..3], then (|o|,|t|,|o&t|) for WT, TC, ET */
DCL_API int dcl_predict_volume(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode,
                       int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host,
                       float* probs_out_dev, uint8_t* labels_out_dev, const uint8_t* target_dev,
                       uint64_t* counts_out_dev, void* stream);

/* Same call with HOST buffers (pinned or pageable): uploads the volume (and target), runs the
 * path, downloads labels (+ probabilities if requested) and the 13 counters.  This is the call
 * the reference's validate_softmax loop maps to (x.cuda() at :133 ... .cpu() at :141). */
DCL_API int dcl_predict_volume_host(dcl_handle* h, const float* vol_host, const int32_t shape[3], int32_t mode,
                            int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host,
                            float* probs_out_host, uint8_t* labels_out_host, const uint8_t* target_host,
                            uint64_t counts_out_host[13], void* stream);

/* BASELINE config 4 ("WT/TC/ET + edge outputs"): dcl_predict_volume in a weighted mode with the six final auxiliary heads
 * (forward()[1] = supervise {01,02,04}, forward()[2] = edge {01,02,04}; cls_wise_former.py:545-546, :585-592) blended with
 * the same weights.  aux_out_dev: (6, 2, X, Y, Z) fp32 in that order; requires cfg.want_aux; UNIFORM / GAUSSIAN only. */
DCL_API int dcl_predict_volume_aux(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode,
                           int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host,
                           float* probs_out_dev, float* aux_out_dev, uint8_t* labels_out_dev,
                           const uint8_t* target_dev, uint64_t* counts_out_dev, void* stream);

/* 8-flip test-time augmentation around the reference tiling: replaces predict_cls.py:180-203:
 *   logit = softmax(T(x)) + sum over the 7 flips f of softmax(flip_f(T(flip_f(x)))),  output = logit / 8,  T = tailor_and_concat,
 * on x[..., :155], flips in the reference's order (none, X, Y, Z, XY, XZ, YZ, XYZ).  keep_scale_host: NULL or 8 x 8 x 16
 * dropout scales in draw order.  probs_out_dev: NULL or (4,240,240,155); labels / target / counts as dcl_predict_volume. */
DCL_API int dcl_predict_volume_tta(dcl_handle* h, const float* vol_dev, const int32_t shape[3], const float* keep_scale_host,
                           float* probs_out_dev, uint8_t* labels_out_dev, const uint8_t* target_dev,
                           uint64_t* counts_out_dev, void* stream);

/* ---- multi-GPU, owner-computes (BASELINE.json configs 3 and 5): the reference runs single-GPU
 * (test_overlap.py:78); this is the sharded form of predict_overlap.tailor_and_concat + the arg-max tail
 * (predict_overlap.py:31-58, :141-153).  Rank r forwards patches [first, first+count) of the z-major plan into slots of
 * ITS OWN memory (dcl_slots_ensure + dcl_forward_patches_to_slots), exports the slot buffer to its peers through CUDA
 * IPC (dcl_ipc_export / dcl_ipc_import: one process per GPU, one node), and after a stream-ordered barrier blends,
 * normalises and labels the x-range [x0, x1) it owns with dcl_gather_finalize_range, whose kernel reads every covering
 * patch's probabilities where they live - local HBM or a peer's over NVLink.  No accumulator is exchanged, every
 * probability crosses NVLink at most once, and the label map is bit-identical to the single-GPU gather form. */
DCL_API int dcl_slots_ensure(dcl_handle* h, int32_t n_slots, void** slots_dev_out);
DCL_API int dcl_ipc_export(const void* dev_ptr, void* handle64_out);
DCL_API int dcl_ipc_import(const void* handle64, void** dev_ptr_out);
DCL_API int dcl_ipc_release(void* dev_ptr);
DCL_API int dcl_forward_patches_to_slots(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode,
                                         int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host,
                                         int32_t first, int32_t count, void* stream);
/* slot_ptrs_host: n_patches device pointers (slot of patch i = 4 x 128^3 floats, local or IPC-mapped); outputs are
 * whole-volume buffers of which only the rows x0 <= x < x1 are written; counts_out_dev += the 13 counters of the range */
DCL_API int dcl_gather_finalize_range(const int32_t shape[3], int32_t mode, int32_t n_patches, const int32_t* starts_host,
                                      const void* const* slot_ptrs_host, int32_t x0, int32_t x1, float* probs_out_dev,
                                      uint8_t* labels_out_dev, const uint8_t* target_dev, uint64_t* counts_out_dev, void* stream);

/* ---- multi-GPU building blocks of the accumulate form (kept for the reduce-scatter variant and the tests): a rank runs
 * only patches [first, first+count) of the plan into its
 * private fp32 accumulator (acc: 4 x X x Y x Zout weighted sums, wsum: X x Y x Zout); the
 * accumulators are then summed across ranks by the caller (NCCL reduce-scatter / all-reduce
 * through torch.distributed) and finalised into labels. */
DCL_API int dcl_accumulate_patches(dcl_handle* h, const float* vol_dev, const int32_t shape[3], int32_t mode,
                           int32_t n_patches, const int32_t* starts_host, const float* keep_scale_host,
                           int32_t first, int32_t count, float* acc_dev, float* wsum_dev, void* stream);
/* labels = argmax_c(acc_c / wsum) over voxels [v0, v0+nvox) of the flattened volume; counters as above. */
DCL_API int dcl_finalize_labels(const float* acc_dev, const float* wsum_dev, int64_t voxels_total, int64_t v0,
                        int64_t nvox, float* probs_out_dev, uint8_t* labels_out_dev,
                        const uint8_t* target_dev, uint64_t* counts_out_dev, void* stream);

/* ---- label export: the savepath / save_format / snapshot block of validate_softmax
 * (predict.py:310-350) and the per-slice pictures / tables of predict_simple.py:186-278 ---- */
/* labels_dev (X,Y,Z) uint8 in {0..3}.  seg_out_dev: NULL or (X,Y,Z) with label 3 written as 4 (predict.py:322-324);
 * seg_nifti_dev: NULL or the same values in NIfTI storage order (x fastest = the array nibabel writes for seg_img);
 * counts_out_dev: NULL or 6 uint64: voxels of label 1, 2, 4, then WT, TC, ET (the verbose print, predict.py:325-328). */
DCL_API int dcl_export_labels(const uint8_t* labels_dev, const int32_t shape[3], uint8_t* seg_out_dev,
                      uint8_t* seg_nifti_dev, uint64_t* counts_out_dev, void* stream);
/* frames_out_dev (Z, X, Y, 3) uint8: frame z = Snapshot_img[:, :, :, z] of predict.py:338-350 with
 * palette[label][rgb]; the reference's schemes are {255,0,0, 0,0,0, 0,255,0, 0,0,255} (predict.py:342-344) and
 * {0,0,0, 250,250,149, 244,130,128, 97,136,200} (predict_simple.py:193-197). */
DCL_API int dcl_snapshot_frames(const uint8_t* labels_dev, const int32_t shape[3], const uint8_t palette[12],
                        uint8_t* frames_out_dev, void* stream);
/* counts_out_dev (Z, 9) uint64: per z slice (|o|,|t|,|o&t|) for WT, TC, ET = the integers behind the per-frame
 * softmax_output_dice of output_excel (predict_simple.py:224-243); Z <= 256. */
DCL_API int dcl_slice_counts(const uint8_t* labels_dev, const uint8_t* target_dev, const int32_t shape[3],
                     uint64_t* counts_out_dev, void* stream);
/* Host-side writers (no device needed).  dcl_write_nifti: NIfTI-1 single file, gzip when the path ends in ".gz";
 * data_host in storage order (x fastest), datatype = NIfTI code (2 uint8, 4 int16, 16 float32, ...); header as
 * nibabel's Nifti1Image(data, None): unit voxels, qform_code = sform_code = 0 (predict.py:329, nibabel itself is
 * absent from this image).  dcl_write_npy_labels: np.save of the int64 arg-max map (predict.py:313-314), byte-identical
 * to numpy 2.x.  dcl_write_png_rgb: one (height, width, 3) frame (imageio.imwrite, predict.py:350). */
DCL_API int dcl_write_nifti(const char* path, const void* data_host, int32_t datatype, const int32_t shape[3]);
DCL_API int dcl_write_npy_labels(const char* path, const uint8_t* labels_host, const int32_t shape[3]);
DCL_API int dcl_write_png_rgb(const char* path, const uint8_t* rgb_host, int32_t height, int32_t width);

/* ---- input side: what data/ClsWiseBraTS128Test.BraDataSet128 (test_overlap.py:14,94-97, not shipped
 * by the reference) hands to predict_overlap.py:132-135 ---- */
DCL_API int dcl_read_nifti_header(const char* path, int32_t shape_out[3], int32_t* datatype_out, float pixdim_out[3]);
/* voxel data as float32 in storage order (x fastest), scl_slope / scl_inter applied; returns the element count */
DCL_API int64_t dcl_read_nifti_f32(const char* path, float* data_out_host, int64_t capacity);
/* modalities_dev (4, Z, Y, X) float32 = four NIfTI arrays as stored (FLAIR, T1ce, T1, T2 in the caller's order) ->
 * vol_out_dev (4, X, Y, z_pad): mask = sum over modalities > 0, per modality (x - mean) / std over the mask (population
 * std), other voxels unchanged, z in [Z, z_pad) zero.  stats_dev: 17 doubles of device scratch; on completion
 * [8] = mask voxels, [9..12] = mean, [13..16] = std. */
DCL_API int dcl_preprocess_volume(const float* modalities_dev, const int32_t shape[3], int32_t z_pad, float* vol_out_dev,
                          double* stats_dev, void* stream);
/* seg_nifti_dev (Z, Y, X) uint8 as stored -> target_out_dev (X, Y, z_pad), optionally 4 -> 3 (predict_overlap.py:150-152) */
DCL_API int dcl_reorder_labels(const uint8_t* seg_nifti_dev, const int32_t shape[3], int32_t z_pad, int32_t map4to3,
                       uint8_t* target_out_dev, void* stream);

/* ---- surface-distance metrics: cal_hausdorff (predict_simple.py:121-144) ->
 * utils/hausdorff.py:86-123 -> medpy.metric.binary.hd95 / hd (medpy is absent: its published algorithm is restated in
 * csrc/hausdorff.cu) for the WT, TC and ET regions of two (X,Y,Z) uint8 label maps in {0..3}.  Exact integer squared
 * distance transform + histogram on the device, numpy's "linear" percentile on the host; 0 for an empty or full mask.
 * The call synchronises the stream.  surface_voxels_out_host: NULL or the number of surface distances per region. */
DCL_API int64_t dcl_hausdorff_workspace_bytes(const int32_t shape[3]);
DCL_API int dcl_hausdorff(const uint8_t* labels_dev, const uint8_t* target_dev, const int32_t shape[3], void* workspace_dev,
                  int64_t workspace_bytes, double hd95_out_host[3], double hd_out_host[3],
                  uint64_t surface_voxels_out_host[3], void* stream);
/* numpy.percentile(a, q) ("linear") of the multiset {sqrt(i) repeated hist[i] times}; pure host code */
DCL_API double dcl_percentile_from_hist(const uint32_t* hist_host, int64_t nbins, double q_percent);

/* ---- introspection used by the parity tests (tests/) and the bench ---- */
/* Copies the named intermediate tensor of the last dcl_forward into out_dev (dense NCDHW / row
 * major).  Returns its element count, or a negative status.  Requires cfg.keep_stages. */
DCL_API int64_t dcl_read_stage(dcl_handle* h, const char* stage, float* out_dev, int64_t cap, void* stream);
/* The 13 top-k index sets (128 int32 each, order: {01,02,04} x {ee,es,ss,se}, fusion). */
DCL_API int dcl_read_topk(dcl_handle* h, int32_t* out_host /* 13*128 */, void* stream);
/* Kernels launched by this handle since creation (the bench's gpu_launches claim). */
DCL_API int64_t dcl_launch_count(const dcl_handle* h);

/* Per-kernel-class device timing (CUDA events on the launching stream, recorded around every launch of
 * the class while enabled).  Classes: 0 = all 3x3x3 convolutions (work = 2*MAC flops), 1 = stitch / accumulate /
 * label kernels (work = algorithmic bytes); bf16 mode also: 2 = rolling conv 16ch@128^3, 3 = rolling conv
 * 32ch@64^3, 4 = slab conv, 5 = im2col GEMM conv (stride 2), 6 = DeUp_Cat, 7 = norm+act+residual, 8 = the token
 * path (select .. couplers .. untokenise, fork to join), 9 = endconv+softmax, 10 = tokenise, 11 = 1x1 convs.
 * dcl_profile_enable(h, 1) starts a new window (clears earlier records); dcl_profile_read synchronises the
 * events and returns the totals of one class over the window. */
DCL_API int dcl_profile_enable(dcl_handle* h, int32_t on);
DCL_API int dcl_profile_read(dcl_handle* h, int32_t cls, double* ms_total, int64_t* launches, double* work_total);

/* ---- single operators, exported for per-op parity tests against torch.nn.functional ---- */
/* y = conv3d(act(norm(cat(x0,x1)))) * out_scale + residual; kernel 3, padding 1, stride 1|2.
 * w is the PyTorch (Cout, C0+C1, 3,3,3) weight, device memory; norm_mean/rstd are per input channel
 * (NULL = identity); act: 0 none, 1 relu, 2 leaky_relu(0.01).  impl: 0 fp32 FFMA, 1 bf16x3, 2 bf16 tcgen05.
 * stats_out: NULL, or 2*cout zero-initialised doubles that receive per-channel (sum, sum of squares) of y
 * (tensor-core impls only: the fused InstanceNorm statistics of the epilogue). */
DCL_API int dcl_op_conv3d_k3(const float* x0, int32_t c0, const float* x1, int32_t c1, const int32_t in_dhw[3],
                     const float* w, const float* bias, int32_t cout, int32_t stride,
                     const float* norm_mean, const float* norm_rstd, int32_t act,
                     const float* residual, float* y, int32_t impl, double* stats_out, void* stream);
DCL_API int dcl_op_instnorm_stats(const float* x, int32_t channels, int64_t spatial, float* mean, float* rstd, void* stream);

/* bench helper (tools/op_time.py): average device time in us of `reps` back-to-back launches of the bf16 kernel the
 * forward picks for a cubic g^3 3x3x3 convolution (mode 1 = fused input norm + residual + statistics) */
DCL_API double dcl_bench_conv(int32_t cin, int32_t cout, int32_t g, int32_t stride, int32_t mode, int32_t reps);

/* bench helper (tools/stitch_time.py): device time in us per volume of the weighted stitch in isolation on a
 * shape[0] x shape[1] x shape[2] volume with the sliding-window plan of `stride`; form 0 = gather (per-patch slots + one
 * gather_finalize launch), form 1 = accumulate (memsets + one accumulate launch per patch + finalize) */
DCL_API double dcl_bench_stitch(const int32_t shape[3], int32_t stride, int32_t gaussian, int32_t form, int32_t reps,
                        int32_t* n_patches_out);

/* debug: 16 %globaltimer stamps (ns) written between the stages of the last bf16 forward when DCL_STAMPS=1 is set:
 * 0 start, 1 encoder done, 2 decoupler + tokenise done, 3 region couplers done, 4 cross-region coupler done,
 * 5/6/7 decoder level starts, 8 decoder done, 9 end */
DCL_API int dcl_debug_stamps(dcl_handle* h, uint64_t* out_host);

/* ---- debug: in-kernel timeline of CTA (0,0) of the tcgen05 kernels (tools/trace_kernel.py) ---- */
DCL_API int dcl_trace_enable(int32_t on);
/* Copies up to `cap` (tag<<32|step, SM clock) pairs recorded since the last read; returns the count. */
DCL_API int64_t dcl_trace_read(int64_t* out_host, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DCL_B200_H */
