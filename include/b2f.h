/* b2f -- C ABI of the B200-native detect / align / embed / match hot path.
 *
 * The reference (Kumar2421/scrfd_arcface_facerecognition) is pure Python and has no FFI layer of its
 * own: its hot path crosses into native code only through third-party wheels (onnxruntime, cv2,
 * scikit-image, numpy/BLAS, qdrant-client).  Each entry point below therefore cites the *reference
 * call site* whose work it replaces; the Python classes in scrfd_arcface_facerecognition_b200/ (and the
 * drop-in `models/`, `utils/` import paths) bind them through ctypes -- see INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success, non-zero on failure (b2f_last_error() has the
 * text); no C++ exceptions cross the boundary; all buffers are caller-owned DEVICE pointers unless
 * named *_host; `stream` is a cudaStream_t passed as void*; nothing allocates after the first call
 * of a given shape.  There is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef B2F_H_
#define B2F_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2F_ABI_VERSION 4

/* dtype codes */
#define B2F_F16 0
#define B2F_BF16 1
#define B2F_F32 2

/* activation codes */
#define B2F_ACT_NONE 0
#define B2F_ACT_RELU 1
#define B2F_ACT_PRELU 2
#define B2F_ACT_SIGMOID 3

int b2f_version(void);
const char* b2f_last_error(void);
/* kernels launched by this library since load (bench.py's gpu_launches) */
long long b2f_launch_count(void);
/* tuning knobs for sweeps: key 0 = smem budget (bytes) for single-N-tile CTAs, 1 = max UMMA N,
 * 2 = persistent conv kernel on/off (off = one CTA per output tile) */
int b2f_set_tuning(int key, int value);

/* ---- a2: aspect-preserving resize + top-left zero letterbox, uint8 out -------------------------
 * replaces cv2.resize + canvas copy at reference models/scrfd.py:135-138 (bit-exact vs cv2
 * INTER_LINEAR 8U, including the exact-2x area path).  frames: [B,H,W,3] u8 BGR; out: [B,in_h,in_w,3]. */
int b2f_letterbox_u8(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w, int in_h,
                     uint8_t* out, void* stream);

/* ---- a2+a3 fused: letterbox + (x-mean)*scale + BGR->RGB, written as NHWC fp16/bf16 with the channel
 * dimension zero-padded to c_pad (>=3) -- the layout the conv kernels consume.
 * replaces reference models/scrfd.py:135-138 + cv2.dnn.blobFromImage at :76-82. */
int b2f_preprocess(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w, int in_h,
                   float mean, float scale, void* out_nhwc, int c_pad, int dtype, void* stream);
/* same, fused with the 3x3 / pad 1 patch extraction of the detector's first convolution: writes
 * [batch][ho][wo][32] patches (k = tap*3 + rgb) that b2f_conv2d consumes as a 1x1 convolution; bit-identical to
 * b2f_preprocess followed by b2f_im2col3x3 (reference models/scrfd.py:76-83, 135-138) */
int b2f_preprocess_patches(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w, int in_h,
                           int stride, float mean, float scale, void* out_patches, int dtype, void* stream);

/* a2 + a3 + the detector's FIRST convolution (3x3 / stride 2 / pad 1, 3 -> cout_p in {16, 32}, bias, optional ReLU) in one
 * kernel: letterboxed, normalised pixels are produced in shared memory, assembled there into the K-major operand tiles of
 * a tcgen05 GEMM (K = 27) and consumed by it, so neither the blob nor a patch tensor reaches HBM (reference models/scrfd.py:76-83, 135-138: resize + blobFromImage + the first Conv node of
 * session.run).  weight [cout_p][32] 16-bit with k = tap * 3 + rgb (27 used), bias [cout_p] f32,
 * out [batch][in_h/2][in_w/2][cout_p] 16-bit. */
int b2f_preprocess_conv1(const uint8_t* frames, int batch, int h, int w, int new_w, int new_h, int in_w, int in_h,
                         float mean, float scale, const void* weight, const float* bias, int cout_p, int act,
                         void* out, int dtype, void* stream);

/* ---- a3 exact: u8 BGR HWC -> fp32 NCHW RGB blob, (float(x)-mean)*scale, bit-exact vs
 * cv2.dnn.blobFromImage(s) (reference models/scrfd.py:76-82, models/arcface.py:44-50). */
int b2f_blob_nchw_f32(const uint8_t* images, int batch, int h, int w, float mean, float scale, float* out,
                      void* stream);

/* ---- a5..a11: anchor decode + threshold + score sort + greedy NMS + max_num selection -----------
 * replaces reference models/scrfd.py:89-119 (forward loop), :142-177 (detect tail), :180-207 (nms) and
 * utils/helpers.py:62-107 (distance2bbox / distance2kps).  Bit-exact float32, numpy-stable tie order. */
typedef struct b2f_det_levels {
  const float* score[3]; /* per level (stride 8,16,32): [B][H/s*W/s][score_ps] , anchors at [0,num_anchors) */
  const float* bbox[3];  /* [B][H/s*W/s][bbox_ps], anchor a at [4a,4a+4)  (stride units) */
  const float* kps[3];   /* [B][H/s*W/s][kps_ps],  anchor a at [10a,10a+10) */
  int score_ps[3], bbox_ps[3], kps_ps[3]; /* floats per pixel (>= 2, 8, 20) */
} b2f_det_levels;

int b2f_decode_nms(const b2f_det_levels* lv, int batch, int in_h, int in_w,
                   const float* det_scale /*[B] device*/, const int* image_hw /*[B][2] device, may be null*/,
                   float conf_thres, float iou_thres, int max_num, int metric /*0=max,1=center-weighted*/,
                   int max_cand, int max_det,
                   float* det /*[B][max_det][5]*/, float* kps /*[B][max_det][10]*/,
                   int* keep_idx /*[B][max_det] index into the score-sorted candidate list, may be null*/,
                   int* counts /*[B][4]: n_det, n_candidates, n_kept_by_nms, overflow flag*/,
                   void* workspace, long long workspace_bytes, void* stream);
long long b2f_decode_nms_workspace(int batch, int max_cand);

/* stand-alone NMS over a (K,5) [x1,y1,x2,y2,score] array; `keep` receives indices into dets in visiting
 * order -- reference SCRFD.nms, models/scrfd.py:180-207.  workspace >= 8 * next_pow2(max(n,32)) bytes. */
int b2f_nms(const float* dets, int n, float iou_thres, int* keep, int* n_keep, void* workspace,
            long long workspace_bytes, void* stream);
/* reference utils/helpers.py:62-83 and :86-107 (float32 add/sub, no clamping: max_shape is never passed) */
int b2f_distance2bbox(const float* points, const float* distance, int n, float* out, void* stream);
int b2f_distance2kps(const float* points, const float* distance, int n, int k2, float* out, void* stream);

/* ---- a12: five-point similarity transform (closed-form 2-D Umeyama, float64) -------------------
 * replaces skimage SimilarityTransform.estimate at reference utils/helpers.py:18-53.
 * landmarks [F][5][2] f32 -> M [F][6] f64 row-major 2x3. */
int b2f_estimate_norm(const float* landmarks, int faces, int image_size, double* m_out, void* stream);

/* ---- a13: cv2.warpAffine(image, M, (size,size), borderValue=0) bit-exact, u8 out ----------------
 * replaces reference utils/helpers.py:58.  frames [B,H,W,3] u8; frame_idx [F]; M [F][6] f64. */
int b2f_warp_affine_u8(const uint8_t* frames, int h, int w, const int* frame_idx, const double* m, int faces,
                       int size, uint8_t* out /*[F][size][size][3]*/, void* stream);

/* ---- a12+a13+a15(blob) fused: landmarks -> aligned, normalised NHWC crop for the embedder --------
 * replaces reference models/arcface.py:54-57 up to session.run (norm_crop_image + blobFromImages).
 * out_nhwc [F][size][size][c_pad] fp16/bf16 RGB, (x-127.5)*float32(1/127.5); crop_u8 optional. */
int b2f_norm_crop(const uint8_t* frames, int h, int w, const int* frame_idx, const float* landmarks, int faces,
                  int size, float mean, float scale, void* out_nhwc, int c_pad, int dtype,
                  uint8_t* crop_u8 /*may be null*/, double* m_out /*may be null*/, void* stream);
/* norm_crop fused with normalisation and the 3x3 / pad 1 / stride 1 patch extraction of ArcFace's first convolution:
 * writes [faces][112][112][32] patches; bit-identical to b2f_norm_crop followed by b2f_im2col3x3
 * (reference utils/helpers.py:18-59, models/arcface.py:44-57) */
int b2f_norm_crop_patches(const uint8_t* frames, int h, int w, const int* frame_idx, const float* landmarks, int faces,
                          int size, float mean, float scale, void* out_patches, int dtype, void* stream);

/* the same crop as 16-byte pixels [faces][112][112][8] (R, G, B, then zeros), the input of b2f_conv2d's stem form
 * (cin_p == 8); bit-identical to b2f_norm_crop with c_pad = 8 */
int b2f_norm_crop_image8(const uint8_t* frames, int h, int w, const int* frame_idx, const float* landmarks, int faces,
                         int size, float mean, float scale, void* out_image8, int dtype, void* stream);

/* ---- a4 / a15: convolution layers of the detector / embedder ------------------------------------
 * replaces onnxruntime session.run at reference models/scrfd.py:83 and models/arcface.py:51.
 * NHWC activations, weights [kh*kw][cout_p][cin_p], fp32 bias table, fused residual + activation. */
typedef struct b2f_conv_desc {
  int n, h, w, cin_p;       /* input  [n][h][w][cin_p]; cin_p == 8 selects the stem form: 3x3 kernel, 16-byte pixels, weight
                             * [10][cout_p][8] with slot = filter tap (ky * 3 + kx) and slot 9 zero, cout_p <= 256 */
  int ho, wo, cout_p;       /* output [n][ho][wo][cout_p] */
  int kh, kw, stride, pad;
  int dtype;                /* activations + weights: B2F_F16 or B2F_BF16 */
  int out_dtype;            /* B2F_F16 / B2F_BF16 / B2F_F32 */
  int act;                  /* B2F_ACT_* */
  int bias_classes;         /* 1, or 9 = per-border-class table for a folded pre-conv BatchNorm shift */
  int res_mode;             /* 0 none, 1 same-size residual, 2 residual is a (res_h,res_w) map upsampled 2x nearest */
  int res_h, res_w;
  int force_kchunk;         /* 0 = auto (64/32/16) */
  int sig_hi;               /* act == SIGMOID: apply to channels [0, sig_hi) only; 0 = all channels */
  const void* in;
  const void* weight;
  const float* bias;        /* [bias_classes][cout_p] */
  const float* slope;       /* [cout_p] for PReLU */
  const void* residual;     /* may be the same buffer as `out` (res_mode 1): the block output replaces its identity input in
                             * place; with act == NONE and a 16-bit output, tiles of <= 128 channels then add through a TMA
                             * reduce-store (the sum is a 16-bit add of the rounded conv result and the stored value) */
  void* out;
  /* optional projection shortcut fused as extra K (ResNet down-sampling blocks: out += conv1x1_stride_s(sc_in)):
   * sc_in [n][sc_h][sc_w][sc_cin_p], sc_weight [1][cout_p][sc_cin_p], no padding, stride sc_stride, same output size;
   * its bias is expected to be folded into `bias`.  NULL = none. */
  const void* sc_in;
  const void* sc_weight;
  int sc_cin_p, sc_stride, sc_h, sc_w;
  /* pool == 1: a 3x3 / stride 2 / pad 1 max-pool of the (ReLU'd, hence non-negative) result is fused into the epilogue:
   * `out` is the POOLED map [n][(ho-1)/2+1][(wo-1)/2+1][cout_p]; every 8x16 conv tile max-reduces its 5x9 partial
   * window maxima into it through TMA (the entry zeroes `out` first).  Needs a 3x3 / stride 1 / pad 1 convolution with
   * act == RELU, a 16-bit output, cout_p % 32 == 0 and no residual (reference graph: Conv-Relu-MaxPool of the SCRFD stem). */
  int pool;
  /* optional split-K workspace (device, >= 8 * n * ho * wo * cout_p * 4 bytes; NULL = never split).  A layer with an fp32
   * output, no activation / residual and >= 128 K chunks per tile (the 7 x 7 x 512 -> 512 embedding layer, reference
   * models/arcface.py:51: 16 work items for 148 SMs) then deals its filter taps to up to 8 work items per tile that store
   * fp32 partial sums here; a second kernel adds them in a fixed order.  Whether a layer splits depends on the layer and on
   * this pointer only, never on n. */
  void* splitk_ws;
  long long splitk_ws_bytes;
} b2f_conv_desc;
int b2f_conv2d(const b2f_conv_desc* desc, void* stream);

/* first layer (cin <= 4, 3x3, pad 1): direct convolution on CUDA cores. weight [3][3][cin_s][cout_p] f32 */
int b2f_stem_conv3x3(const void* in, int n, int h, int w, int cin_s /*stored channels*/, int stride,
                     const float* weight, const float* bias, const float* slope, int act, int cout_p,
                     int dtype, void* out, void* stream);
/* 3x3 pad-1 patches of a 4-channel (RGB0) NHWC image -> [n][ho][wo][32] with k = tap*3 + channel (27 used):
 * turns the first layer into a 1x1 convolution for the tensor-core kernel */
int b2f_im2col3x3(const void* in, int n, int h, int w, int stride, int ho, int wo, int dtype, void* out, void* stream);
/* depthwise kxk (+bias, activation). weight [k*k][c_p] f32 */
int b2f_dwconv(const void* in, int n, int h, int w, int c_p, int k, int stride, int pad, const float* weight,
               const float* bias, const float* slope, int act, int dtype, void* out, void* stream);
/* pooling: mode 0 = max (pad with -inf), 1 = average (ceil_mode, count_include_pad=0) */
int b2f_pool(const void* in, int n, int h, int w, int c_p, int k, int stride, int pad, int mode, int ho, int wo,
             int dtype, void* out, void* stream);

/* out = act(a*scale[c] + shift[c] + b): fallback for BatchNormalization / Add / activation nodes that no
 * convolution absorbed (scale/shift/b may be null). */
int b2f_eltwise(const void* a, const void* b, long long pixels, int c_p, const float* scale, const float* shift,
                const float* slope, int act, int dtype, void* out, void* stream);

/* ---- a17 / a22: cosine similarity pieces -----------------------------------------------------------
 * rows -> unit rows (fp32 and/or 16-bit copies) + norms; replaces reference utils/helpers.py:110-123 norms
 * and duplicate.py:1491-1496. */
int b2f_l2norm_rows(const float* x, long long rows, int dim, float* out_f32 /*may be null*/,
                    void* out_16 /*may be null*/, int dtype, float* norms /*may be null*/, void* stream);
/* compute_similarity for `pairs` vector pairs: dot / (|a| |b|) in fp32 (reference utils/helpers.py:110-123) */
int b2f_cosine_pairs(const float* a, const float* b, int pairs, int dim, float* out, void* stream);

/* ---- a18 / a19: gallery matching -------------------------------------------------------------------
 * coarse pass: tcgen05 GEMM of queries x gallery^T with a running top-k epilogue (never materialises
 * Q x G); replaces the per-target Python loop at reference main.py:136-142 and the Qdrant scan behind
 * qdrant_manager.py:164-170.  part_* are [Q][2*n_splits][topk] (each CTA reports its two column halves). */
int b2f_match_partial(const void* queries, int q, const void* gallery, long long g, int dim, int dtype,
                      const float* row_scale, const float* col_scale, int topk, int n_splits,
                      float* part_score, int* part_idx, void* stream);
/* same, with `keep` <= topk candidates actually tracked per list (the lists keep their topk stride, unused slots are empty):
 * a top-1 query only needs a short list for the exact re-score, and the list length is what the epilogue costs */
int b2f_match_partial_keep(const void* queries, int q, const void* gallery, long long g, int dim, int dtype,
                           const float* row_scale, const float* col_scale, int topk, int keep, int n_splits,
                           float* part_score, int* part_idx, void* stream);
/* prefix ("causal") variant: query row i, whose global index is causal_base + i, is matched only against the gallery rows
 * before it -- the online clustering loop's "best earlier person" (reference duplicate.py:1853-1855) for all rows at once */
int b2f_match_partial_causal(const void* queries, int q, const void* gallery, long long g, int dim, int dtype, int topk,
                             int keep, int n_splits, long long causal_base, float* part_score, int* part_idx, void* stream);
/* exact fp32 scores of one unit query against all stored unit rows (a Qdrant search with k = every row, reference
 * duplicate.py:2757-2766): out[r] = <rows[r], query> */
int b2f_rows_dot(const float* rows, long long n, int dim, const float* query, float* out, void* stream);
int b2f_match_splits(long long g, int want);
/* the n_splits b2f_match_partial should be called with for q queries against g gallery rows: ranges per query tile that
 * keep every SM (pair) evenly busy.  part_* must then hold [q][2*n_splits][topk] entries. */
int b2f_match_plan(int q, long long g);
/* merge pass: exact fp32 re-score of the coarse candidates, sort (score desc, index asc), threshold.
 * q_f32 / g_f32 are unit-norm fp32 rows; idx_base is added to indices (gallery shard offset). */
int b2f_match_merge(const float* part_score, const int* part_idx, int q, int n_cand, const float* q_f32,
                    const float* g_f32, int dim, int topk, float threshold, int strict_gt, long long idx_base,
                    float* out_score /*[Q][topk]*/, long long* out_idx /*[Q][topk], -1 = none*/, void* stream);

/* sharded top-1 exchange (gallery rows split over GPUs, per-shard results merged by score desc, index asc):
 * (score, global index >= 0 or -1) <-> one int64 key whose signed maximum over the shards IS that merge, so the
 * exchange is a single MAX all-reduce of Q keys (an empty slot is INT64_MIN).  Global indices must be < 2^32 - 1. */
int b2f_topk_pack_keys(const float* score, const long long* idx, long long n, long long* keys, void* stream);
int b2f_topk_unpack_keys(const long long* keys, long long n, float* score, long long* idx, void* stream);

/* ---- a21: duplicate merge (greedy one-hop leader clustering in ascending id order) ------------------
 * replaces the N Qdrant searches of reference duplicate.py:2726-2797. */
int b2f_pairs_threshold(const void* emb16, int n, int dim, int dtype, int row_begin, int row_end, float threshold,
                        const float* emb_f32 /*unit rows for exact re-check, may be null*/,
                        long long* pairs /*[max_pairs] (i<<32|j), j>i*/, long long max_pairs,
                        unsigned long long* pair_count, void* stream);
int b2f_cluster_resolve(const long long* pairs_sorted, long long n_pairs, int n, int* leader, void* stream);

/* ---- 8f rank 3: overlay drawing on device-resident frames -----------------------------------------
 * replaces the cv2.rectangle / cv2.line / cv2.putText calls of reference utils/helpers.py:126-179 (draw_bbox,
 * draw_bbox_info; called per face at reference main.py:144-148) for frames that stay in HBM.  The caller lowers those
 * calls to draw commands (overlay.py): kind 0 = fill the inclusive rectangle [x0, x1] x [y0, y1] (clipped to the frame),
 * kind 1 = paint `bgr` wherever the 1-byte-per-pixel mask at masks + mask_off (x1 columns, y1 rows, row-major) is
 * non-zero, mask pixel (0, 0) landing on frame pixel (x0, y0).  Commands are grouped (one group = the commands of one
 * face, which share a colour); group g holds commands [group_cmds[g], group_cmds[g + 1]) and frame f owns groups
 * [frame_groups[f], frame_groups[f + 1]), painted in that order (a later face paints over an earlier one, as in the
 * reference loop).  frames: [batch][h][w][3] u8 BGR, modified in place.  All pointers are device pointers. */
typedef struct b2f_draw_cmd {
  int kind;
  int x0, y0, x1, y1;
  unsigned int bgr;      /* b | g << 8 | r << 16 */
  int mask_off;
  int reserved;
} b2f_draw_cmd;
int b2f_draw_overlay(uint8_t* frames, int batch, int h, int w, const b2f_draw_cmd* cmds, const int* frame_groups,
                     const int* group_cmds, const uint8_t* masks, void* stream);

/* ---- debug: one TMA box -> raw shared-memory image (tests the tensor-map conventions) -------------- */
int b2f_debug_tma_probe(const void* src, const long long* dims4, const int* box4, const int* estr4,
                        int swizzle_bytes, const int* coords4, unsigned char* out_dev, int out_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2F_H_ */
