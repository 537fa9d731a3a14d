/* qlidar.h -- C ABI of the B200-native (sm_100a) quantized sparse-3D-conv backbone path.
 *
 * The reference (BiboyQG/Quantization-on-3D-Object-Detection) has NO C/FFI boundary of its own for this path:
 * it is Python that calls two third-party packages, spconv 2.x (pybind `core_cc`, not vendored) and
 * pytorch_quantization.  Each entry point below therefore cites the reference *call site* whose work it
 * replaces (paths relative to the reference root).  The Python mirror of the reference's module API
 * (quantization-on-3d-object-detection_b200/qlidar) binds exactly these symbols through ctypes; the stub a
 * maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - no allocation, no synchronisation, no global state: the caller owns all buffers and the stream;
 *  - row counts live on the device (`n_dev`, int32) so a whole backbone forward is sync-free and can be
 *    captured in a CUDA graph; `n_cap` is the host-known capacity used to size grids and buffers; a NULL
 *    `n_dev` means "exactly n_cap rows";
 *  - coords are int32 [b, z, y, x] rows; the linear key ((b*D+z)*H+y)*W+x must fit in 32 bits;
 *  - return value: QL_OK or a negative QL_ERR_* code; nothing throws.
 */
#ifndef QLIDAR_H_
#define QLIDAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ql_stream_t; /* cudaStream_t */

enum {
    QL_OK = 0,
    QL_ERR_INVALID = -1,        /* bad argument (null pointer, unsupported channel count, ...) */
    QL_ERR_CUDA = -2,           /* a CUDA runtime call or launch failed; see ql_last_cuda_error() */
    QL_ERR_GRID_TOO_LARGE = -3, /* B*D*H*W does not fit the 32-bit key */
    QL_ERR_WORKSPACE = -4,      /* workspace too small */
    QL_ERR_UNSUPPORTED = -5
};

enum { QL_F16 = 0, QL_F32 = 1, QL_S8 = 2, QL_S32 = 3 };

/* quantize_rows modes */
enum {
    QL_Q_CODES_PER_TENSOR = 0, /* int8 codes, one scale for the tensor   (QConvNd cw=False, quant/quant.py:28-32) */
    QL_Q_FAKE_PER_CHANNEL = 1, /* fp16 fake-quant, amax per input channel (QConvNd cw=True,  quant/quant.py:21-26) */
    QL_Q_FAKE_PER_TENSOR = 2,  /* fp16 fake-quant, one amax */
    QL_Q_FAKE_PER_ROW = 3      /* fp16 fake-quant, amax per voxel row     (GQConv3d, quant/quant_conv3d.py:112-131) */
};

int ql_abi_version(void);
const char* ql_error_string(int code);
const char* ql_last_cuda_error(void);
int ql_num_sms_on_device(void);

/* ---- coordinate hash (replaces [EXT] spconv's hash table behind SubMConv3d/SparseConv3d.forward,
 *      call sites pcdet/models/backbones_3d/spconv_backbone.py:12-17; SPCONV_USE_DIRECT_TABLE=False,
 *      pcdet/utils/spconv_utils.py:4-5).  Table = uint64 slots {key:32 | value:32}, capacity a power of two. */
int64_t ql_hash_capacity(int64_t max_entries);
int ql_hash_build(const int32_t* coords, int64_t n_cap, const int32_t* n_dev,
                  int32_t B, int32_t D, int32_t H, int32_t W,
                  uint64_t* table, int64_t table_cap, ql_stream_t stream);

/* ---- voxelization + mean VFE (replaces VoxelGeneratorWrapper.generate -> [EXT] Point2VoxelCPU3d,
 *      pcdet/datasets/processor/data_processor.py:45-61,151-153; collate_batch's batch column,
 *      pcdet/datasets/dataset.py:237-244; MeanVFE.forward, pcdet/models/backbones_3d/vfe/mean_vfe.py:25-29;
 *      with max_pts_per_voxel == 0: DynamicMeanVFE.forward, .../vfe/dynamic_mean_vfe.py:53-72).
 *      points: [n_points, point_stride] fp32; column 0 is the batch index when has_batch_col != 0, then x,y,z,
 *      then the remaining features (n_feat counts x,y,z).  Voxels are numbered in first-touch order over the
 *      point array (== spconv's CPU voxelizer); a voxel keeps its first `max_pts_per_voxel` points; voxels
 *      numbered >= max_voxels are dropped.  max_voxels_per_frame > 0 additionally keeps only the first that many voxels of
 *      every frame (the reference voxelises frame by frame with MAX_NUMBER_OF_VOXELS each, data_processor.py:151-153); it needs
 *      the points of a frame to be contiguous and the frames in ascending order (collate_batch's layout).  Outputs: out_feats [max_voxels, out_feat_stride >= n_feat] fp32 (mean; the pad
 *      columns are zero -- a stride of 8 lets ql_stem_conv fetch a row with one 256-bit load), out_coords
 *      [max_voxels, 4] int32, out_npts [max_voxels] int32, n_voxels_dev int32[2] = {rows kept, voxels found before
 *      the cap}, and the hash table coords -> row. */
size_t ql_voxelize_workspace_bytes(int64_t max_points, int64_t max_voxels, int32_t n_feat, int32_t max_pts_per_voxel);
int ql_voxelize_mean(const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                     const float* range_min_xyz_host, const float* voxel_size_xyz_host, const int32_t* grid_xyz_host,
                     int32_t batch_size, int32_t max_pts_per_voxel, int64_t max_voxels, int64_t max_voxels_per_frame,
                     float* out_feats, int32_t out_feat_stride, int32_t* out_coords, int32_t* out_npts, int32_t* n_voxels_dev,
                     uint64_t* table, int64_t table_cap, void* workspace, size_t workspace_bytes, ql_stream_t stream);
/* The two halves of ql_voxelize_mean (same arguments and workspace): ql_voxelize_coords runs the hash insert and the
 * numbering passes -- out_coords and n_voxels_dev are final after it -- and ql_voxelize_features, stream-ordered after it, the point
 * selection and the means (out_feats, out_npts, table values).  Work that needs only coordinates (rulebooks) can overlap it. */
int ql_voxelize_coords(const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                     const float* range_min_xyz_host, const float* voxel_size_xyz_host, const int32_t* grid_xyz_host,
                     int32_t batch_size, int32_t max_pts_per_voxel, int64_t max_voxels, int64_t max_voxels_per_frame,
                     float* out_feats, int32_t out_feat_stride, int32_t* out_coords, int32_t* out_npts, int32_t* n_voxels_dev,
                     uint64_t* table, int64_t table_cap, void* workspace, size_t workspace_bytes, ql_stream_t stream);
int ql_voxelize_features(const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                     const float* range_min_xyz_host, const float* voxel_size_xyz_host, const int32_t* grid_xyz_host,
                     int32_t batch_size, int32_t max_pts_per_voxel, int64_t max_voxels, int64_t max_voxels_per_frame,
                     float* out_feats, int32_t out_feat_stride, int32_t* out_coords, int32_t* out_npts, int32_t* n_voxels_dev,
                     uint64_t* table, int64_t table_cap, void* workspace, size_t workspace_bytes, ql_stream_t stream);
/* MeanVFE on an already voxelized (V,T,F) tensor (mean_vfe.py:25-29); num_points is float32 when
 * num_points_is_float (pcdet/models/__init__.py:36 casts everything to float) else int32. */
int ql_mean_vfe(const float* voxels, const void* num_points, int32_t num_points_is_float, int64_t V, int32_t T, int32_t F,
                float* out_feats, ql_stream_t stream);

/* ---- rulebook / indice pairs (replaces [EXT] spconv indice-pair generation inside
 *      SubMConv3d/SparseConv3d.forward; keys at spconv_backbone.py:194-231).
 *      Layout produced: nbr[tile][k][128] int32, tile = row/128: the input row feeding output row
 *      tile*128+r through kernel offset k = (kz*KH+ky)*KW+kx, or -1.  ksize/stride/pad are zyx triples.
 *      tile_kmask (nullable): uint32 [tiles][ql_rulebook_mask_words(kvol)], bit k set iff some row of the tile has a
 *      neighbour through offset k.  WITH a mask the rulebook is COMPACT: only a tile's live slabs are written, first in
 *      the tile's block and in ascending k -- nbr[tile][j][128], j = number of set mask bits below k; the block keeps its
 *      kvol*512-byte stride and the slabs past popc(mask) are undefined.  The conv kernels take (nbr, tile_kmask) pairs
 *      in exactly this form (tile_kmask == NULL: the dense layout, every offset visited).
 *      Submanifold: outputs == inputs (same rows); neighbours are found by open-addressing hash probes on `table`.
 *      Strided: active output sites are numbered in ascending order of the linear key ((b*Do+z)*Ho+y)*Wo+x (spconv
 *      leaves the order implementation-defined); out_table receives out coords -> row for the layers that follow.
 *      n_out_dev is int32[2] = {rows kept (<= n_out_cap), active output sites found}; [1] > [0] means the caller's
 *      capacity overflowed (the sites with the largest keys are dropped and never referenced). */
int64_t ql_rulebook_num_tiles(int64_t n_out_cap);
int32_t ql_rulebook_mask_words(int32_t kvol);
int ql_rulebook_subm(const int32_t* coords, int64_t n_cap, const int32_t* n_dev,
                     int32_t B, int32_t D, int32_t H, int32_t W, const int32_t* ksize_host,
                     const uint64_t* table, int64_t table_cap, int32_t* nbr_out, uint32_t* tile_kmask, ql_stream_t stream);
size_t ql_rulebook_strided_workspace_bytes(int32_t B, int32_t D, int32_t H, int32_t W,
                                           const int32_t* ksize_host, const int32_t* stride_host, const int32_t* pad_host);
int ql_rulebook_strided(const int32_t* in_coords, int64_t n_in_cap, const int32_t* n_in_dev,
                        int32_t B, int32_t D, int32_t H, int32_t W,
                        const int32_t* ksize_host, const int32_t* stride_host, const int32_t* pad_host,
                        int32_t* out_coords, int64_t n_out_cap, int32_t* n_out_dev,
                        uint64_t* out_table, int64_t out_table_cap, int32_t* nbr_out, uint32_t* tile_kmask,
                        void* workspace, size_t workspace_bytes, ql_stream_t stream);

/* ---- rank index of a key-sorted site list (no reference counterpart: it replaces the hash probes spconv does for the
 *      layers that FOLLOW a strided conv on the same stage).  ql_rulebook_strided leaves, in its workspace, a bitmap over
 *      the output grid's cells and an exclusive popcount prefix per 32-cell word; for the rows it produced (ascending key
 *      order) row(key) = word_prefix[key/32] + popc(bitmap[key/32] & ((1 << key%32) - 1)).  `out_table` of
 *      ql_rulebook_strided may be NULL when every consumer of the stage uses the rank index instead of the hash.
 *      ql_rulebook_strided_index (host only) returns the two device pointers inside `workspace`; they stay valid until the
 *      workspace is reused.  ql_rulebook_subm_ranked == ql_rulebook_subm on such a stage. */
int ql_rulebook_strided_index(int32_t B, int32_t D, int32_t H, int32_t W,
                              const int32_t* ksize_host, const int32_t* stride_host, const int32_t* pad_host,
                              void* workspace, const uint32_t** bitmap, const uint32_t** word_prefix, int64_t* n_words);
/* ql_rulebook_strided for a KEY-SORTED input stage: the pairs are taken from the output side through the input stage's rank
 * index (in_bitmap / in_word_prefix over the INPUT grid B,D,H,W), so no -1 fill / scatter / mask pass is needed.  Same
 * outputs and workspace as ql_rulebook_strided (whose rank index of the OUTPUT stage it also leaves behind); no hash. */
int ql_rulebook_strided_ranked(const int32_t* in_coords, int64_t n_in_cap, const int32_t* n_in_dev,
                               int32_t B, int32_t D, int32_t H, int32_t W,
                               const int32_t* ksize_host, const int32_t* stride_host, const int32_t* pad_host,
                               const uint32_t* in_bitmap, const uint32_t* in_word_prefix,
                               int32_t* out_coords, int64_t n_out_cap, int32_t* n_out_dev,
                               int32_t* nbr_out, uint32_t* tile_kmask,
                               void* workspace, size_t workspace_bytes, ql_stream_t stream);
/* Renumber a list of DISTINCT sites (any order, e.g. the voxeliser's first-touch order == the reference's CPU voxeliser order,
 * data_processor.py:151-153) by ascending linear key: out_coords [n] sorted, n_out_dev = (n, n), src_row[r] = input row of
 * output row r (optional), and -- optional payload -- rows_out[r] = rows_in[src_row[r]] (rows of row_bytes = 16 m bytes).
 * Leaves the stage's rank index in `workspace` in the layout of a 1x1x1 stride-1 ql_rulebook_strided over (B, D, H, W):
 * ql_rulebook_strided_workspace_bytes / ql_rulebook_strided_index with ksize = stride = {1,1,1}, pad = {0,0,0}. */
int ql_renumber_by_key(const int32_t* in_coords, int64_t n_cap, const int32_t* n_dev,
                       int32_t B, int32_t D, int32_t H, int32_t W,
                       int32_t* out_coords, int32_t* n_out_dev, int32_t* src_row,
                       const void* rows_in, void* rows_out, int32_t row_bytes,
                       void* workspace, size_t workspace_bytes, ql_stream_t stream);
int ql_rulebook_subm_ranked(const int32_t* coords, int64_t n_cap, const int32_t* n_dev,
                            int32_t B, int32_t D, int32_t H, int32_t W, const int32_t* ksize_host,
                            const uint32_t* bitmap, const uint32_t* word_prefix,
                            int32_t* nbr_out, uint32_t* tile_kmask, ql_stream_t stream);

/* ---- GROUPED submanifold rulebook (the counterpart of spconv's MaskImplicitGemm mask argsort, the GPU default the
 *      reference runs with -- ConvAlgo.MaskImplicitGemm, tools/demo.ipynb:255).  The output rows are binned by a 9-bit line
 *      key (which of the 3 x 3 kernel x-lines around the site hold any active cell), so that the rows of one 128-row MMA
 *      tile share their live kernel offsets and the conv kernel skips the rest.  row_perm_out [tiles * 128]: tile slot ->
 *      output row (-1 = padding of the last tile); nbr_out / tile_kmask are laid out by SLOT and must be consumed through
 *      ql_spconv_mma_rows / ql_stem_conv_rows with the same row_perm.  Feature and coordinate arrays keep their row order:
 *      results are bit-identical to the ungrouped rulebook's.  The slot order inside a bin is not deterministic. */
size_t ql_rulebook_group_workspace_bytes(int64_t n_cap);
int ql_rulebook_subm_ranked_grouped(const int32_t* coords, int64_t n_cap, const int32_t* n_dev,
                                    int32_t B, int32_t D, int32_t H, int32_t W, const int32_t* ksize_host,
                                    const uint32_t* bitmap, const uint32_t* word_prefix,
                                    int32_t* nbr_out, uint32_t* tile_kmask, int32_t* row_perm_out,
                                    void* workspace, size_t workspace_bytes, ql_stream_t stream);

/* ---- implicit gather-GEMM-scatter sparse conv on tcgen05 tensor cores (replaces QConvNd.forward ->
 *      [EXT] spconv conv forward, quant/quant.py:36-58, plus the BatchNorm1d/ReLU/residual that follow it in
 *      post_act_block / SparseBasicBlock, spconv_backbone.py:8-27,51-67).
 *      feats: [n_in, c_in] fp16 (kind::f16, fp32 accumulate) or int8 codes (kind::i8, int32 accumulate); the gathered
 *      rows go global -> registers -> tensor memory (A operand read from TMEM), never through shared memory.
 *      ZERO-ROW CONTRACT: the row in front of the features -- the c_in * elem_size bytes at feats - c_in * elem_size --
 *      must be readable and all zero.  A missing neighbour is rulebook index -1, and the gather reads row -1 for it
 *      without a predicate or a select (round 2: the kernel was bound by instruction issue, profiles/r02_conv_ablation.md).
 *      qlidar.ops.zero_led_rows allocates such buffers; qlidar.ops.spconv_mma copies any other tensor into one.
 *      nbr / tile_kmask: a rulebook from ql_rulebook_subm / ql_rulebook_strided (compact when tile_kmask is given);
 *      tile_kmask may be NULL (dense rulebook, every offset is visited); kvol <= 128.
 *      w_packed: per-output-channel int8 codes (as fp16 exact integers for the f16 kind) in the shared-memory
 *      image built by ql_pack_weights_host.  Epilogue: y = acc * (scale[oc] * (act_scale_dev ? *act_scale_dev : 1))
 *      + shift[oc] (+ residual) ; optional ReLU; written as out_dtype (QL_F16/QL_F32) -- or the raw
 *      accumulators when out_dtype == QL_S32 (int32 for the i8 kind, fp32 bits for the f16 kind).
 *      Optional second output out_q = clamp(rint(y * out_qscale[oc]), -127, 127) int8 (static re-quantization
 *      for the next layer) and absmax[oc] (atomic max of |y|, fp32, caller zero-initialises). */
size_t ql_packed_weight_bytes(int32_t c_in, int32_t c_out, int32_t kvol, int32_t elem_dtype);
int ql_pack_weights_host(const void* w_host, int32_t elem_dtype, int32_t c_in, int32_t c_out, int32_t kvol,
                         void* packed_host);
int ql_spconv_mma(const void* feats, int32_t in_dtype, const int32_t* nbr, const uint32_t* tile_kmask,
                  int64_t n_out_cap, const int32_t* n_out_dev,
                  int32_t c_in, int32_t c_out, int32_t kvol, const void* w_packed,
                  const float* scale, const float* shift, const float* act_scale_dev,
                  const void* residual_f16, int32_t relu,
                  void* out, int32_t out_dtype, int8_t* out_q, const float* out_qscale, float* absmax,
                  ql_stream_t stream);
/* the same through a grouped rulebook: row_perm [tiles * 128] maps a tile slot to the output row it computes (out,
 * out_q, residual are indexed by ROW; nbr / tile_kmask by slot); row_perm == NULL is ql_spconv_mma.
 * ql_spconv_weights_streamed (host only): 1 when the kernel streams this layer's packed weights per unit because they do
 * not fit in shared memory (fp16 C >= 64, int8 C >= 128) -- those launches are paced by the L2 -> SM weight stream. */
int32_t ql_spconv_weights_streamed(int32_t c_in, int32_t c_out, int32_t kvol, int32_t elem_dtype);
int ql_spconv_mma_rows(const void* feats, int32_t in_dtype, const int32_t* nbr, const uint32_t* tile_kmask,
                       const int32_t* row_perm, int64_t n_out_cap, const int32_t* n_out_dev,
                       int32_t c_in, int32_t c_out, int32_t kvol, const void* w_packed,
                       const float* scale, const float* shift, const float* act_scale_dev,
                       const void* residual_f16, int32_t relu,
                       void* out, int32_t out_dtype, int8_t* out_q, const float* out_qscale, float* absmax,
                       ql_stream_t stream);

/* ---- row permutation out[r] = in[src_row[r]] (rows of row_bytes = 16 m bytes, 16-byte aligned; src_row < 0 -> zero row).
 *      No reference counterpart: the engine renumbers the voxeliser's first-touch-ordered voxels (== the reference's CPU
 *      voxeliser order, data_processor.py:151-153) into ascending-key order with a 1x1x1 ql_rulebook_strided build, whose
 *      rulebook is src_row, so that stage 1 is indexed by rank (bitmap + prefix) like the stages strided convs produce. */
int ql_permute_rows(const void* in, void* out, int32_t row_bytes, const int32_t* src_row, int64_t n_cap,
                    const int32_t* n_dev, ql_stream_t stream);

/* ---- fp32 SIMT stem conv for the un-quantized conv_input (C_in = 4/5 raw point features; quant_centerpoint.py
 *      backbone_no_list = ['backbone_3d.conv_input.0'], :24-26).  feats rows are feat_stride floats apart;
 *      w: [kvol][c_in][c_out] fp32. */
int ql_stem_conv(const float* feats, int32_t feat_stride, int32_t c_in, const int32_t* nbr, const uint32_t* tile_kmask, int64_t n_out_cap, const int32_t* n_out_dev,
                 int32_t c_out, int32_t kvol, const float* w, const float* scale, const float* shift, int32_t relu,
                 void* out, int32_t out_dtype, float* absmax, ql_stream_t stream);

/* ---- activation quantizer ([EXT] pytorch_quantization TensorQuantizer as configured at quant/quant.py:14-32).
 *      absmax_cols: absmax[c] = max(absmax[c], max_rows |x[:,c]|).  quantize_rows: see QL_Q_* modes; `smooth`
 *      (nullable, [c]) divides x first (SmoothQuant, quant/smoothquant.py:72-79); act_scale_out[0] receives the
 *      de-quantization scale amax/bound for QL_Q_CODES_PER_TENSOR. */
int ql_absmax_cols(const void* x, int32_t dtype, int64_t n_cap, const int32_t* n_dev, int32_t c, float* absmax,
                   ql_stream_t stream);
int ql_quantize_rows(const void* x, int32_t in_dtype, int64_t n_cap, const int32_t* n_dev, int32_t c,
                     const float* absmax, const float* smooth, int32_t bits, int32_t mode,
                     void* out, float* act_scale_out, ql_stream_t stream);

/* ---- SmoothQuant weight preparation on the device (quant/smoothquant.py:69-82 formula, per input channel, as intended for the
 *      sparse convs by quant/quant_conv3d.py:141-236): smooth[ic] = act_absmax[ic]^alpha / w_ic_absmax[ic]^(1-alpha) (zeros -> 1),
 *      w'[oc,k,ic] = w * smooth[ic], per-output-channel int8 codes of w' written into the packed image ql_spconv_mma reads
 *      (same bytes as ql_pack_weights_host of those codes) and scale_out[oc] = amax(w'[oc]) / 127 * (bn_scale ? bn_scale[oc] : 1).
 *      w: [c_out][kvol][c_in] fp32 (device); w_ic_absmax: [c_in] max over (oc, k) of |w|; act_absmax: [c_in] (dynamic: the
 *      producing layer's epilogue; static: calibrated).  The activations are then quantised with ql_quantize_rows(smooth). */
int ql_sq_prepare_weights(const float* w, const float* w_ic_absmax, const float* act_absmax, float alpha,
                          int32_t c_in, int32_t c_out, int32_t kvol, const float* bn_scale,
                          float* smooth_out, int8_t* packed_out, float* scale_out, ql_stream_t stream);

/* ---- dense SmoothQuant wrappers (quant/smoothquant.py:38-99 SQConv2d.forward and its 1-D / transposed / linear siblings;
 *      quant/SQSubM2d.py:22-91): per-COLUMN statistics and int8 codes of F.unfold(x) without materialising it in floating point.
 *      x: dense NCHW fp16/fp32; kernel / stride / pad / dilation are (h, w) pairs (host); columns in F.unfold order
 *      c*kh*kw + ky*kw + kx.  ql_unfold_absmax: absmax_cols[col] = max(absmax_cols[col], max over windows |x|) (caller zeroes).
 *      ql_unfold_quantize: out[m][col] int8 = clamp(rint(x / smooth_cols[col] * (bound / amax_t))), m = (b, oy, ox), rows
 *      col_stride bytes apart (>= C*kh*kw, multiple of 16, pad columns 0), amax_t = max_col absmax_cols[col] / smooth_cols[col]
 *      (the per-tensor input quantiser of quantize.py:59-76); scales_out = {bound / amax_t, amax_t / bound}. */
int ql_unfold_absmax(const void* x, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W,
                     const int32_t* kernel_hw, const int32_t* stride_hw, const int32_t* pad_hw, const int32_t* dil_hw,
                     float* absmax_cols, ql_stream_t stream);
int ql_unfold_quantize(const void* x, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W,
                       const int32_t* kernel_hw, const int32_t* stride_hw, const int32_t* pad_hw, const int32_t* dil_hw,
                       const float* absmax_cols, const float* smooth_cols, int32_t bits, int32_t col_stride,
                       int8_t* out, float* scales_out, ql_stream_t stream);

/* ---- BEV hand-off (replaces HeightCompression.forward -> [EXT] SparseConvTensor.dense(),
 *      pcdet/models/backbones_2d/map_to_bev/height_compression.py:20-24): out[b, c*D+d, y, x], zero filled. */
size_t ql_bev_densify_workspace_bytes(int32_t B, int32_t D, int32_t H, int32_t W);
int ql_bev_densify(const void* feats, int32_t in_dtype, int32_t c, const uint64_t* table, int64_t table_cap,
                   int32_t B, int32_t D, int32_t H, int32_t W, void* out, int32_t out_dtype,
                   void* workspace, size_t workspace_bytes, ql_stream_t stream);

/* same hand-off for a key-sorted stage, cell -> row taken from the stage's rank index (rows >= min(n_cap, *n_dev) are absent) */
int ql_bev_densify_ranked(const void* feats, int32_t in_dtype, int32_t c, const uint32_t* bitmap, const uint32_t* word_prefix,
                          int64_t n_cap, const int32_t* n_dev, int32_t B, int32_t D, int32_t H, int32_t W, void* out,
                          int32_t out_dtype, void* workspace, size_t workspace_bytes, ql_stream_t stream);

/* ---- VoxelNeXt 2-D merge (replaces VoxelResBackBone8xVoxelNeXt.bev_out, spconv_backbone_voxelnext.py:149-164:
 *      indices[:, [0, 2, 3]] -> torch.unique(dim=0, return_inverse) -> index_add_).  coords: [n, 4] int32 [b, z, y, x] (z is
 *      dropped; the caller has already scaled the coarser stages onto the target grid, :194-197).  Outputs: out_coords
 *      [n_out_cap, 3] int32 [b, y, x] in ascending (b, y, x) order, out_feats [n_out_cap, c] (QL_F32 or QL_F16; c % 4 == 0) =
 *      the sum of the rows that fall on each site (fp32 atomics), n_out_dev int32[2] = {rows kept, sites found}. */
size_t ql_bev_merge2d_workspace_bytes(int32_t B, int32_t H, int32_t W, int64_t n_out_cap, int32_t c, int32_t out_dtype);
int ql_bev_merge2d(const void* feats, int32_t in_dtype, int32_t c, const int32_t* coords, int64_t n_cap, const int32_t* n_dev,
                   int32_t B, int32_t H, int32_t W, void* out_feats, int32_t out_dtype, int32_t* out_coords,
                   int64_t n_out_cap, int32_t* n_out_dev, void* workspace, size_t workspace_bytes, ql_stream_t stream);

/* the same for SEVERAL site lists merged in one pass -- VoxelResBackBone8xVoxelNeXt.forward's concat of stages 4, 5, 6
 * (spconv_backbone_voxelnext.py:194-199): segment i contributes rows feats[i] / coords[i] (counts n_dev[i], capacity n_cap[i]) with
 * y and x multiplied by coord_scale[i] (1, 2, 4: `indices[:, 1:] *= 2` / `*= 4`; z is dropped, so its scaling does not matter).
 * The pointer arrays are HOST arrays of device pointers.  out_coord_cols = 3: [b, y, x] (torch.unique order); 4: [b, 0, y, x],
 * the form the rulebook entry points take for the 2-D tail. */
int ql_bev_merge2d_multi(int32_t n_seg, const void* const* feats, int32_t in_dtype, int32_t c, const int32_t* const* coords,
                         const int64_t* n_cap, const int32_t* const* n_dev, const int32_t* coord_scale,
                         int32_t B, int32_t H, int32_t W, void* out_feats, int32_t out_dtype, int32_t* out_coords, int32_t out_coord_cols,
                         int64_t n_out_cap, int32_t* n_out_dev, void* workspace, size_t workspace_bytes, ql_stream_t stream);

/* ---- key-sorted voxelisation straight from the points (engine front end; no reference counterpart: the reference's voxeliser is
 *      first-touch ordered, spconv's order downstream is implementation-defined, and the engine works in ascending-key order).
 *      Same point layout / range / voxel size / grid arguments as ql_voxelize_mean.  The voxel id of a point is the rank of its cell
 *      in the stage-1 bitmap, so the call also leaves the stage's rank index in rank_workspace (layout of
 *      ql_rulebook_strided_index(B, grid_z + 1, grid_y, grid_x, {1,1,1}, {1,1,1}, {0,0,0})).  No per-frame voxel cap in this order:
 *      frame_counts [B] (optional) reports the voxels found per frame so that the caller can fall back to ql_voxelize_mean.
 *      ql_voxelize_sorted_features (same arguments and workspace, stream-ordered after the coordinates half; max_pts >= 1) writes the
 *      means of each voxel's first max_pts points (by index; bit-identical to ql_voxelize_mean's) and the point counts. */
size_t ql_voxelize_sorted_workspace_bytes(int64_t max_points, int64_t max_voxels, int32_t max_pts);
int ql_voxelize_sorted_coords(const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                              const float* range_min, const float* vsize, const int32_t* grid_xyz, int32_t batch_size,
                              int64_t max_voxels, int32_t* out_coords, int32_t* n_voxels_dev, int32_t* frame_counts,
                              void* rank_workspace, size_t rank_workspace_bytes, void* workspace, size_t workspace_bytes,
                              ql_stream_t stream);
int ql_voxelize_sorted_features(const float* points, int64_t n_points, int32_t point_stride, int32_t has_batch_col, int32_t n_feat,
                                const float* range_min, const float* vsize, const int32_t* grid_xyz, int32_t batch_size,
                                int32_t max_pts, int64_t max_voxels, const int32_t* n_voxels_dev, float* out_feats,
                                int32_t out_feat_stride, int32_t* out_npts, void* workspace, size_t workspace_bytes,
                                ql_stream_t stream);

/* ---- CenterHead post-processing, all on the device (replaces CenterHead.generate_predicted_boxes,
 *      pcdet/models/dense_heads/center_head.py:297-365 -> centernet_utils._topk / decode_bbox_from_heatmap
 *      (model_utils/centernet_utils.py:155-241) -> model_nms_utils.class_agnostic_nms (model_nms_utils.py:6-25) ->
 *      iou3d_nms_utils.nms_gpu (ops/iou3d_nms/iou3d_nms_utils.py:120-135) -> nms_kernel + the HOST sweep of iou3d_nms.cpp
 *      (ops/iou3d_nms/src/iou3d_nms_kernel.cu:295-339, iou3d_nms.cpp:137-183)).
 *
 *  ql_centerhead_decode: one head.  Maps are fp32 NCHW device arrays: hm [B,C,H,W] LOGITS (sigmoid is applied here), center
 *      [B,2,H,W], center_z [B,1,H,W], dim [B,3,H,W] (log sizes; exp is applied here), rot [B,2,H,W] (cos, sin), vel [B,2,H,W] or
 *      NULL, iou [B,1,H,W] or NULL.  K = MAX_OBJ_PER_SAMPLE (<= 1024).  voxel_size_xy / pc_min_xy / center_limit_range
 *      (POST_CENTER_LIMIT_RANGE, 6 floats) are HOST arrays.  score_thresh < 0 = no threshold.  class_map: device int32 [C] or
 *      NULL (class_id_mapping_each_head).  Outputs per frame, in descending score order, masked rows removed (the reference's
 *      boolean-mask indexing): out_boxes [B,K,7 or 9] (x, y, z, dx, dy, dz, heading[, vx, vy]), out_scores [B,K], out_labels
 *      [B,K] (mapped class, 0-based), out_iou [B,K] ((iou + 1) / 2) when iou is given, out_count [B]. */
size_t ql_centerhead_decode_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W);
int ql_centerhead_decode(const float* hm, const float* center, const float* center_z, const float* dim, const float* rot,
                         const float* vel, const float* iou, int32_t B, int32_t C, int32_t H, int32_t W, int32_t K,
                         float feature_map_stride, const float* voxel_size_xy, const float* pc_min_xy,
                         const float* center_limit_range, float score_thresh, const int32_t* class_map, float* out_boxes,
                         float* out_scores, int32_t* out_labels, float* out_iou, int32_t* out_count, void* workspace,
                         size_t workspace_bytes, ql_stream_t stream);

/* ---- VoxelNeXt sparse head post-processing (replaces VoxelNeXtHead.generate_predicted_boxes,
 *      pcdet/models/dense_heads/voxelnext_head.py:418-488 -> centernet_utils._topk_1d / gather_feat_idx /
 *      decode_bbox_from_voxels_nuscenes (model_utils/centernet_utils.py:243-354) and rotate_class_specific_nms_iou (:308-331)).
 *
 *  ql_voxelhead_decode: ql_centerhead_decode over a SPARSE head: per-voxel row-major fp32 arrays hm [N,C] logits, center [N,2],
 *      center_z [N,1], dim [N,3] log sizes, rot [N,2] (cos, sin), vel [N,2] / NULL, iou [N,1] / NULL (raw head output: (iou+1)/2
 *      clamped to [0,1] is applied here), indices_byx int32 [N,3] = the 2-D sparse tensor's (batch, y, x), n_dev = device row count
 *      or NULL (= n_cap).  Per frame: top-K over its (voxel, class) scores, decode, range / score mask.  Outputs as
 *      ql_centerhead_decode.
 *  ql_voxelhead_class_split: the IoU-branch re-scoring: per (class, frame) the frame's decoded boxes of that class with score
 *      score^(1-r[c]) * iou^(r[c]), sorted by it; out_* are [num_class, B, K(, box_dim)], out_counts [num_class, B] -- one
 *      ql_nms_rotated call per class (its own threshold / pre / post sizes) finishes rotate_class_specific_nms_iou. */
size_t ql_voxelhead_decode_workspace_bytes(int32_t B, int32_t C, int64_t n_cap);
int ql_voxelhead_decode(const float* hm, const float* center, const float* center_z, const float* dim, const float* rot,
                        const float* vel, const float* iou, const int32_t* indices_byx, int64_t n_cap, const int32_t* n_dev,
                        int32_t B, int32_t C, int32_t K, float feature_map_stride, const float* voxel_size_xy,
                        const float* pc_min_xy, const float* center_limit_range, float score_thresh, const int32_t* class_map,
                        float* out_boxes, float* out_scores, int32_t* out_labels, float* out_iou, int32_t* out_count,
                        void* workspace, size_t workspace_bytes, ql_stream_t stream);
int ql_voxelhead_class_split(const float* boxes, int32_t box_dim, const float* scores, const int32_t* labels, const float* ious,
                             const int32_t* counts, int32_t B, int32_t K, int32_t num_class, const float* rectifier_dev,
                             float* out_boxes, float* out_scores, int32_t* out_labels, int32_t* out_counts, ql_stream_t stream);

/*  ql_nms_rotated: greedy rotated-BEV-IoU NMS per frame.  boxes [B, n_cap, box_stride] fp32, each frame's first counts[b] rows
 *      (device int32 [B]; NULL = n_cap) valid and ALREADY in descending score order (ql_centerhead_decode's order; nms_gpu sorts
 *      first, iou3d_nms_utils.py:127-131).  The first min(count, pre_max) boxes take part (NMS_PRE_MAXSIZE); a box is suppressed
 *      when its IoU with a kept earlier box is > thresh; at most post_max are kept (NMS_POST_MAXSIZE).  Outputs: keep
 *      [B, post_max] indices into the frame's rows (-1 padded), keep_count [B]; optional gathers out_boxes [B, post_max, box_dim],
 *      out_scores / out_labels [B, post_max] from scores / labels [B, n_cap] (labels get label_offset added: the reference's
 *      final "+ 1").  iou_out (optional, tests): [B, n_cap, n_cap], upper-triangle pairwise IoU.  n_cap <= 2048. */
size_t ql_nms_rotated_workspace_bytes(int32_t B, int32_t n_cap);
int ql_nms_rotated(const float* boxes, int32_t box_stride, int32_t box_dim, const float* scores, const int32_t* labels,
                   const int32_t* counts, int32_t B, int32_t n_cap, float thresh, int32_t pre_max, int32_t post_max,
                   int32_t label_offset, int32_t* keep, int32_t* keep_count, float* out_boxes, float* out_scores,
                   int32_t* out_labels, float* iou_out, void* workspace, size_t workspace_bytes, ql_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* QLIDAR_H_ */
