/* mci_b200.h -- C ABI of the B200-native slice-pair interpolation path.
 *
 * Drop-in boundary for itk::MorphologicalContourInterpolator's hot path
 * (reference: include/itkMorphologicalContourInterpolator.hxx, "X" below; .h = "H").
 * Plain C, plain pointers and sizes; no torch/ITK types.  Every function returns an
 * int status (MCI_OK = 0), never throws and never calls back into the host.  One host
 * thread per context.  There is NO CPU fallback behind this ABI: if no CUDA device is
 * usable, mci_create fails.
 *
 * Two layers are exported from the same shared library (libmci_b200.so):
 *   1. phase-level entry points (mci_upload_volume ... mci_download_volume): thin
 *      wrappers around batched sm_100a kernels; these are what a host scheduler calls;
 *   2. the host scheduler itself (mci_filter_*): C++ host code that restates
 *      GenerateData / InterpolateAlong / InterpolateBetweenTwo's pair bookkeeping
 *      (X:1240-1637) and calls ONLY the phase-level functions of layer 1.
 * INTEGRATION.md shows the itk::ImageToImageFilter subclass whose GenerateData()
 * forwards to layer 2, and the ctypes binding used by the Python package.
 */
#ifndef MCI_B200_H
#define MCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCI_OK 0
#define MCI_ERR_CUDA 1        /* a CUDA call failed; see mci_last_error */
#define MCI_ERR_ARG 2         /* bad argument */
#define MCI_ERR_STATE 3       /* phase called out of order */
#define MCI_ERR_CAPACITY 4    /* a device-side arena / table overflowed */
#define MCI_ERR_PIXELTYPE 5   /* component count or distance bin does not fit the pixel type (X:431-437) */
#define MCI_ERR_NO_DEVICE 6

/* pixel types of the label volume (H: TImage::PixelType).  GPU scope: 3-D u8 / u16 / i16. */
#define MCI_U8 0
#define MCI_U16 1
#define MCI_I16 2

typedef struct mci_ctx mci_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int mci_create(mci_ctx ** ctx, int device);
void mci_destroy(mci_ctx * ctx);
const char * mci_last_error(const mci_ctx * ctx);
/* page-locked host buffers (callers that want full PCIe bandwidth for upload/download) */
void * mci_host_alloc(size_t bytes);
void mci_host_free(void * p);
/* number of kernel launches issued by this context since creation (bench: gpu_launches) */
int64_t mci_launch_count(const mci_ctx * ctx);
/* CUDA stream the context launches on (cudaStream_t as void*), for callers that time with events */
void * mci_stream(const mci_ctx * ctx);
int mci_synchronize(mci_ctx * ctx);
/* How the host thread waits for the bulk volume copies (100+ MB, milliseconds): 0 (default) = spin, lowest latency
 * for one volume at a time; 1 = sleep on a blocking event, for when several contexts / ranks share the host cores
 * (mci_filter_run_batch sets it on its contexts).  Short waits (kernel phases) always spin. */
int mci_set_wait_mode(mci_ctx * ctx, int sleep_on_bulk_copies);
/* per-kernel device times, CUDA events recorded on the launching stream around every launch while enabled */
typedef struct
{
  char name[32];
  int64_t launches;
  double total_ms;
} mci_kernel_time;
int mci_profile_enable(mci_ctx * ctx, int on); /* 0 off, 1 every launch, 2 only the HBM-streaming kernel k_copy_flag */
int mci_profile_read(mci_ctx * ctx, mci_kernel_time * out, int32_t cap, int32_t * n);

/* ---- volume ------------------------------------------------------------------------------- */
/* Replaces GetInput()/AllocateOutputs() (X:1541-1565).  dims = {X, Y, Z}, X fastest.
 * LIMITS: every dimension must be in [1, 8192] (the bound the device-side tables, tile grids and 32-bit distance
 * arithmetic are sized for: a slice diagonal squared stays below 2^28), a slice may hold at most 65 535 connected components of one label (the
 * component images are uint16; 255 for uint8 volumes, as in the reference, whose component image has the volume's
 * pixel type), and labels are the non-zero values of the pixel type (negative int16 labels included).  Violations
 * return MCI_ERR_ARG / MCI_ERR_PIXELTYPE; nothing is truncated silently.
 * mci_upload_volume copies host voxels to the device; mci_attach_device_volume adopts
 * caller-owned device buffers (input is read-only; output must hold X*Y*Z pixels). */
int mci_upload_volume(mci_ctx * ctx, const void * host_voxels, const int64_t dims[3], int pixel_type);
int mci_attach_device_volume(mci_ctx * ctx, const void * dev_in, void * dev_out, const int64_t dims[3], int pixel_type);
/* Output label volume (X:1623-1636 already applied) to host / leave on device. */
int mci_download_volume(mci_ctx * ctx, void * host_voxels);
void * mci_device_output(mci_ctx * ctx);
int mci_volume_dims(const mci_ctx * ctx, int64_t dims[3], int * pixel_type);

/* ---- slice detection: DetermineSliceOrientations (X:128-201) --------------------------------
 * One streaming pass over the volume: per-label bounding boxes (X:148-159) and, for every
 * voxel whose two neighbours along exactly one axis are 0 and whose neighbours along the
 * other two equal it, slice index -> m_LabeledSlices[axis][label] (X:161-197), restricted
 * to `axis` unless axis == -1 (X:193).  The same pass initialises the output volume with a
 * copy of the input (the final overwrite of X:1623-1636 folded forward; interpolation only
 * writes where the input is 0). */
typedef struct
{
  int64_t label;
  int32_t bbox_index[3];
  int32_t bbox_size[3];
} mci_label_info;
typedef struct
{
  int32_t axis;
  int32_t slice;
  int64_t label;
} mci_labeled_slice;
int mci_detect_slices(mci_ctx * ctx, int axis, int32_t * n_labels, int32_t * n_labeled_slices);
int mci_get_detection(mci_ctx * ctx, mci_label_info * labels, mci_labeled_slice * slices);

/* ---- connected components: RegionedConnectedComponents (X:1224-1238) ------------------------
 * For a batch of annotated slices: extract the label's bounding-box slice, binarise
 * (== label), label 2-D components with ids 1..n in raster order of first pixel (ITK
 * ConnectedComponentImageFilter); fully_connected = UseBallStructuringElement (H:156).
 * Slice coordinates (u, v) are the two remaining volume axes in increasing order (X:681-693).
 * Also accumulates what Centroid (X:1101-1134) needs per component. */
typedef struct
{
  int32_t axis;
  int32_t slice;
  int64_t label;
  int32_t u0, v0, w, h; /* 2-D region of the component image */
} mci_slice_desc;
typedef struct
{
  int64_t count;
  int64_t sum_u, sum_v; /* sums of absolute slice indices */
  int32_t min_u, min_v, max_u, max_v;
} mci_component_info;
int mci_label_slices(mci_ctx * ctx, int32_t n, const mci_slice_desc * slices, int fully_connected,
                     int32_t * n_components /* [n] */);
/* all components of all slices, concatenated in slice order (sum of n_components entries) */
int mci_get_components(mci_ctx * ctx, mci_component_info * out);
/* one component image, as uint16 ids, row-major w*h (parity tests) */
int mci_get_component_image(mci_ctx * ctx, int32_t slice_index, uint16_t * ids);

/* ---- component matching: pair scan of InterpolateBetweenTwo (X:1250-1272) -------------------
 * Lock-step scan of the two component images of every segment; emits each distinct
 * (iId, jId) != (0, 0) once, in no particular order.  The set algebra of X:1274-1446
 * stays on the host. */
typedef struct
{
  int32_t slice_i, slice_j; /* indices into the mci_label_slices batch */
} mci_segment_desc;
typedef struct
{
  int32_t segment;
  uint16_t i_id, j_id;
} mci_pair;
int mci_match_pairs(mci_ctx * ctx, int32_t n_segments, const mci_segment_desc * segments, int32_t * n_pairs);
int mci_get_pairs(mci_ctx * ctx, mci_pair * pairs);
/* mci_label_slices + mci_match_pairs launched back to back with a single host round trip (what the scheduler uses) */
int mci_label_and_match(mci_ctx * ctx, int32_t n, const mci_slice_desc * slices, int fully_connected, int32_t n_segments,
                        const mci_segment_desc * segments, int32_t * n_components /* [n] */, int32_t * n_pairs);

/* ---- interpolation jobs ---------------------------------------------------------------------
 * One job = one Align (X:1136-1222) followed by
 *   MCI_JOB_ONE_TO_ONE : Interpolate1to1(..., recursive=false)            X:1323-1324
 *   MCI_JOB_ONE_TO_N   : Interpolate1toN (split, then n 1-to-1 trees)     X:1352-1353,1384-1385,1427-1428
 *   MCI_JOB_EXTRAPOLATE: Extrapolate (phantom point)                      X:1288,1296 -> X:203-251
 * `slice_a` holds the single component `a_id` (the "i" side of the call); `slice_b` holds the
 * n_b components listed at b_ids[b_ids_offset ...] in ascending id order (ignored for
 * EXTRAPOLATE).  pos_a / pos_b are the slice positions along the axis as passed to the
 * reference call (so M-to-1 and (0,b) extrapolations arrive with the roles already swapped). */
#define MCI_JOB_ONE_TO_ONE 0
#define MCI_JOB_ONE_TO_N 1
#define MCI_JOB_EXTRAPOLATE 2
typedef struct
{
  int32_t kind;
  int32_t slice_a, a_id;
  int32_t slice_b, n_b, b_ids_offset;
  int32_t pos_a, pos_b;
} mci_job;
typedef struct
{
  int32_t heuristic_alignment;    /* H:105-113 */
  int32_t use_distance_transform; /* H:115-126 */
  int32_t use_ball;               /* H:150-165 */
  int32_t min_align_iters;        /* X:87   (2^Dim = 8)   */
  int32_t max_align_iters;        /* X:89   (6^Dim = 216) */
} mci_params;
/* Runs Align, the 1-to-N splits and every level of the Interpolate1to1 recursion for all jobs,
 * writing mid slices into the device output with the max rule (X:754-763).  Asynchronous; errors
 * raised on the device surface at the next mci_synchronize / mci_download_volume. */
int mci_interpolate(mci_ctx * ctx, int32_t n_jobs, const mci_job * jobs, const int32_t * b_ids, int32_t n_b_ids,
                    const mci_params * params);
/* translations chosen by Align, 2 ints per job (parity tests) */
int mci_get_translations(mci_ctx * ctx, int32_t * t /* [2*n_jobs] */);
/* per job: translations popped, scoring waves, SM clock cycles, kind*1000 + n_b (diagnostics) */
int mci_get_align_stats(mci_ctx * ctx, int32_t * out /* [4*n_jobs] */);

/* dst[i] = max(dst[i], src[i]) over n pixels of the context's pixel type: the write rule (X:754-763) applied to
 * partial outputs of sharded runs (device pointers; runs on the context's stream) */
int mci_combine_max(mci_ctx * ctx, void * dev_dst, const void * dev_src, int64_t n_pixels, int pixel_type);

/* Multi-GPU combine by shapes instead of by voxels: after a (sharded) run, the mid slices it produced are available
 * as a list of 48-byte records {x0,y0,w,h,pitch,pad,word offset; axis,label,mid,pad} plus their bit-packed rows
 * (device pointers into the context, valid until the next run).  mci_apply_masks replays such a list -- typically
 * another GPU's, received over NCCL -- into this context's output volume with the write rule of X:754-763. */
int mci_get_emitted(mci_ctx * ctx, void ** recs_dev, int64_t * n_recs, void ** words_dev, int64_t * n_words);
/* packs the list into a caller-owned device buffer: records at dst, bit rows at dst + words_offset_bytes */
int mci_copy_emitted(mci_ctx * ctx, void * dst_dev, int64_t words_offset_bytes);
int mci_apply_masks(mci_ctx * ctx, const void * recs_dev, int64_t n_recs, const void * words_dev, int64_t n_words);

/* ---- label smoothing applied to the interpolator's output by the reference's example --------
 * itk::MedianImageFilter<TImage,TImage> with SetRadius(radius) (reference examples/mciExample.cxx:29-39): box
 * neighbourhood (2r+1)^3, ZeroFluxNeumann boundary, element n/2 of the sorted neighbourhood.  radius in [0, 3].
 * _resident: the context's device output volume (e.g. what mci_filter_run_resident left there) is replaced by its
 * smoothed version, asynchronously on the context's stream; follow with mci_download_volume.
 * _run: host in -> host out (upload, filter, download; synchronises). */
int mci_median_filter_resident(mci_ctx * ctx, int radius);
int mci_median_filter_run(mci_ctx * ctx, const void * host_in, void * host_out, const int64_t dims[3], int pixel_type,
                          int radius);

/* ---- host scheduler (layer 2): the filter itself --------------------------------------------
 * Mirrors the setters of H:85-165 and GenerateData (X:1559-1637). */
typedef struct
{
  int64_t label;                      /* SetLabel, 0 = all            H:85-92   */
  int32_t axis;                       /* SetAxis, -1 = all            H:94-101  */
  int32_t heuristic_alignment;        /* default 1                    H:231     */
  int32_t use_distance_transform;     /* default 1                    H:232     */
  int32_t use_ball_structuring_element; /* default 0                  H:233     */
  int32_t use_custom_slice_positions; /* default 0                    H:234     */
  int32_t use_extrapolation;          /* default 1                    H:235     */
  /* multi-GPU: this run interpolates only segments k with k % shard_count == shard_index, k counting the
   * segments of InterpolateAlong (X:1456-1521) in (axis, label, slice) order.  Defaults 0 / 1 = everything. */
  int32_t shard_index, shard_count;
} mci_filter_options;
typedef struct
{
  int64_t n_labels, n_annotated_slices, n_segments, n_pairs, n_jobs, n_components;
  int64_t n_launches;
  double ms_upload, ms_detect, ms_components, ms_pairs, ms_interpolate, ms_download, ms_total;
} mci_filter_stats;
void mci_filter_default_options(mci_filter_options * opt);
/* custom_slices: n_custom triplets (axis, label, slice) used when use_custom_slice_positions (H:188-206) */
int mci_filter_run(mci_ctx * ctx, const mci_filter_options * opt, const void * host_in, void * host_out,
                   const int64_t dims[3], int pixel_type, const int64_t * custom_slices, int64_t n_custom,
                   mci_filter_stats * stats);
/* same, input already on the device (mci_upload_volume / mci_attach_device_volume) and output left there */
int mci_filter_run_resident(mci_ctx * ctx, const mci_filter_options * opt, const int64_t * custom_slices,
                            int64_t n_custom, mci_filter_stats * stats);
/* Slab-sharded run (multi-GPU, one process per GPU; replaces the per-thread segment loop of X:1523-1537 by a partition
 * of the segments over GPUs by the slab their mid slices fall in).  The context's volume (mci_attach_device_volume /
 * mci_upload_volume) is a contiguous range of the whole volume along the interpolation axis: this rank's slab plus
 * the annotated slices bounding the segments that reach into it.  `labels` (bounding boxes, X:148-159) and `slices`
 * (m_LabeledSlices, X:161-197) are the detection tables MERGED OVER ALL RANKS (each rank runs mci_detect_slices on
 * its slab plus one halo slice per side and contributes the entries of its own slices), shifted into the coordinates
 * of the range; slices outside the range are left out.  Only segments with a mid slice in [keep_lo, keep_hi) of the
 * range are interpolated.  The output buffer must already hold a copy of the input over [keep_lo, keep_hi)
 * (mci_detect_slices does that for the range it ran on; mci_copy_input_range for anything else).  Synchronises. */
int mci_filter_run_with_detection(mci_ctx * ctx, const mci_filter_options * opt, const mci_label_info * labels,
                                  int32_t n_labels, const mci_labeled_slice * slices, int32_t n_slices, int32_t keep_lo,
                                  int32_t keep_hi, mci_filter_stats * stats);
/* out[first_voxel ... first_voxel + n_voxels) = in[...] on the context's stream (device to device) */
int mci_copy_input_range(mci_ctx * ctx, int64_t first_voxel, int64_t n_voxels);
/* Throughput mode for a series of independent label volumes (same dims and pixel type): volume k goes through
 * ctxs[k % ...] -- n_ctx contexts, each with its own stream, driven by one host thread each -- so the H2D copy of
 * one volume, the kernels of another and the D2H copy of a third overlap (PCIe is full duplex; a single mci_filter_run
 * leaves each direction idle most of the time).  Results are those of n_volumes mci_filter_run calls.  host_in[k] /
 * host_out[k] should be pinned (mci_host_alloc) for the copies to be asynchronous.  stats: [n_volumes] or NULL.
 * Returns the first non-zero status of any volume (its message is in that context's mci_last_error). */
int mci_filter_run_batch(mci_ctx * const * ctxs, int32_t n_ctx, const mci_filter_options * opt, const void * const * host_in,
                         void * const * host_out, int32_t n_volumes, const int64_t dims[3], int pixel_type,
                         mci_filter_stats * stats);
/* slice sets found by the last run / by mci_filter_detect, as (axis,label,slice) triplets:
 * GetLabeledSliceIndices (H:209-224) */
int mci_filter_detect(mci_ctx * ctx, const mci_filter_options * opt, int64_t * n_triplets);
int mci_filter_get_labeled_slices(mci_ctx * ctx, int64_t * triplets /* [3*n] */);
/* the pair bookkeeping of one segment (X:1274-1446) as a pure host function: needs no device (host-logic tests).
 * use_extrapolation: bit 0 = the option; bit 1 set = take the literal std::set algebra even for pair sets the scheduler
 * handles on its allocation-free path (extrapolations and 1-to-1 pairs only), so that tests can compare the two */
int mci_filter_schedule_pairs(const mci_pair * pairs, int32_t n_pairs, int use_extrapolation, mci_job * jobs, int32_t cap_jobs,
                              int32_t * n_jobs, int32_t * b_ids, int32_t cap_b_ids, int32_t * n_b_ids);
/* drops the scheduler's per-context state (called by mci_destroy) */
void mci_filter_forget(mci_ctx * ctx);

#ifdef __cplusplus
}
#endif
#endif /* MCI_B200_H */
