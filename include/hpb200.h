/*
 * hpb200.h -- C ABI of libhpb200.so: the B200-native (sm_100a) MegaPose / CosyPose
 * render-and-compare hot path (batched rasteriser, perspective crop, image-space pose update,
 * segmented top-K) behind happypose's PosePredictor / PoseEstimator / Panda3dBatchRenderer API.
 *
 * Conventions
 *   - every entry point returns 0 on success and a negative HPB_E* code on failure;
 *     hpb_last_error() returns a thread-local, NUL-terminated description of the last failure.
 *   - pointers named *_dev are DEVICE pointers on the context's device, *_host are host pointers.
 *   - matrices are row-major float32: TCO = 16 floats (4x4), K = 9 floats (3x3).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Nothing
 *     synchronises the host unless stated; work is ordered on `stream`.
 *   - images are planar CHW float32 with a caller-chosen batch stride (in elements), so a caller can
 *     have the kernels write straight into a slice of a larger [b, C_total, h, w] network input.
 *
 * Reference interfaces replaced (paths relative to the happypose repo root):
 *   hpb_render            happypose/toolbox/renderer/panda3d_batch_renderer.py:194-286 (render) with
 *                         :144-192 (make_scene_data), :62-125 (worker_loop) and
 *                         panda3d_scene_renderer.py:320-390 (render_scene)
 *   hpb_mesh_upload       panda3d_scene_renderer.py:206-219 (get_object_node: load, scale, hpr)
 *   hpb_crop              happypose/pose_estimators/megapose/models/pose_rigid.py:199-277 (crop_inputs):
 *                         toolbox/lib3d/camera_geometry.py:40-67,70-122 and
 *                         toolbox/lib3d/cropping.py:27-75,113-152,155-197 (roi_align, sampling_ratio 4)
 *   hpb_crop_boxes        pose_rigid.py:279-337 (compute_crops_multiview, return_crops=False)
 *   hpb_normalize_T       toolbox/lib3d/transform_ops.py:107-120 + rotations.py:22-36
 *   hpb_pose_update       pose_rigid.py:339-350 -> toolbox/lib3d/cosypose_ops.py:34-62;
 *                         cosypose/lib3d/cosypose_ops.py:18-42 (apply_imagespace_predictions)
 *   hpb_tco_init          toolbox/lib3d/cosypose_ops.py:159-181,184-238,241-283
 *   hpb_multiview         toolbox/lib3d/multiview.py:28-92,166-251
 *   hpb_topk_segmented    toolbox/utils/tensor_collection.py:201-230 (filter_top_pose_estimates)
 *   hpb_normalize_depth   pose_rigid.py:455-544
 *   hpb_icp_points        megapose/inference/icp_refiner.py:138-176,271-289 + refiner_utils.py (compute_masks)
 */
#ifndef HPB200_H
#define HPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPB_VERSION 100

/* error codes */
#define HPB_OK 0
#define HPB_EINVAL (-1)  /* bad argument (shape, NULL, range)            */
#define HPB_ECUDA (-2)   /* a CUDA runtime call failed                    */
#define HPB_ENOMEM (-3)  /* host or device allocation failed              */
#define HPB_ENOTFOUND (-4) /* unknown mesh id                             */

/* hpb_render output selection (rgb is always produced by the reference; here it is a flag too) */
#define HPB_RENDER_RGB 1u
#define HPB_RENDER_NORMALS 2u
#define HPB_RENDER_DEPTH 4u
#define HPB_RENDER_MASK 8u

/* hpb_pose_update variants */
#define HPB_POSE_MEGAPOSE 0      /* pose_update_with_reference_point, 9-D output (6-D rot + vxvyvz) */
#define HPB_POSE_COSYPOSE_6D 1   /* apply_imagespace_predictions, 9-D output                        */
#define HPB_POSE_COSYPOSE_QUAT 2 /* apply_imagespace_predictions, 7-D output (xyzw quaternion)       */

/* hpb_tco_init variants */
#define HPB_TCO_INIT_AUTODEPTH_WITH_R 0 /* TCO_init_from_boxes_autodepth_with_R (MegaPose coarse) */
#define HPB_TCO_INIT_ZUP_AUTODEPTH 1    /* TCO_init_from_boxes_zup_autodepth (CosyPose)           */
#define HPB_TCO_INIT_FROM_BOXES 2       /* TCO_init_from_boxes(z_range) (CosyPose)                */

/* hpb_normalize_depth kinds (PosePredictor.depth_normalization_type) */
#define HPB_DEPTH_NORM_NONE 0
#define HPB_DEPTH_NORM_TCR_SCALE 1
#define HPB_DEPTH_NORM_TCR_SCALE_CLAMP_CENTER 2
#define HPB_DEPTH_NORM_TCR_CENTER_CLAMP 3

/* hpb_multiview types (PosePredictor.multiview_type) */
#define HPB_MV_TCO_FRONT_1VIEW 0
#define HPB_MV_TCO_FRONT_3VIEWS 1
#define HPB_MV_SPHERE_26VIEWS 2

typedef struct hpb_ctx hpb_ctx;

int hpb_version(void);
const char *hpb_last_error(void);

/* One context per (process, device).  Owns mesh/texture device buffers and the rasteriser workspace. */
int hpb_create(int device, hpb_ctx **out);
int hpb_destroy(hpb_ctx *ctx);

/* Number of kernels this library has launched since the context was created (bench.py gpu_launches). */
int64_t hpb_launch_count(const hpb_ctx *ctx);

/*
 * Workspaces (rasteriser visibility buffer / vertex scratch, the crop's interleaved frame copy, top-K scratch, the
 * device mesh table) are grown on demand by the launch calls and are NEVER freed or moved while the context lives: a
 * buffer that has to grow is replaced and the old one retired until hpb_destroy, so kernel parameters baked into a
 * captured CUDA graph stay valid.  hpb_workspace_epoch() counts such replacements (a caller may re-capture its graphs
 * when it changes).  Growth cannot happen while `stream` is being captured (allocation is illegal there; the launch
 * returns HPB_EINVAL): either run the call once eagerly first, or size everything up front with hpb_reserve():
 *   render_h, render_w  largest render resolution               (0 = skip)
 *   frame_pixels        n_im * H * W of the largest crop source (0 = skip)
 *   topk_rows/groups    largest hpb_topk_segmented input        (0 = skip)
 * The shared scratch is used by one stream at a time: when consecutive launches come from different (non-capturing)
 * streams the library orders them with an event, so two streams never race on it.
 */
int64_t hpb_workspace_epoch(const hpb_ctx *ctx);

/*
 * Number of rendered scenes, since the context was created (or since the last call with reset != 0), in which at least one
 * mesh vertex lay in front of the near plane.  The rasteriser DROPS the triangles of such vertices instead of clipping them
 * against the plane as OpenGL does (stated deviation); this counter lets a caller (and the parity tests) check
 * that a workload never gets there: MegaPose / CosyPose place objects at 0.3 .. 2 m with z_near = 0.1 m.  Synchronises.
 */
int hpb_raster_clipped_scenes(hpb_ctx *ctx, int64_t *count, int reset);
int hpb_reserve(hpb_ctx *ctx, int render_h, int render_w, int64_t frame_pixels, int64_t topk_rows, int64_t topk_groups);

/*
 * Upload one mesh (HOST pointers; synchronous, not on the hot path).
 *   verts_xyz  [n_verts,3] vertex positions already in METRES and with any ypr offset applied
 *   normals    [n_verts,3] unit normals, or NULL (area-weighted smooth normals are generated)
 *   uv         [n_verts,2] texture coordinates (v up, GL convention), or NULL
 *   vcolor     [n_verts,4] RGBA8 vertex colours, or NULL
 *   faces      [n_faces,3] int32 vertex indices (triangles only)
 *   tex        [tex_h,tex_w,tex_c] uint8 texture (tex_c = 3 or 4), or NULL; a box-filter mip chain is
 *              built on the device
 * Returns the mesh id in *mesh_id (dense, starting at 0).
 */
int hpb_mesh_upload(hpb_ctx *ctx, const float *verts_xyz, const float *normals, const float *uv,
                    const uint8_t *vcolor, int64_t n_verts, const int32_t *faces, int64_t n_faces,
                    const uint8_t *tex, int tex_h, int tex_w, int tex_c, int32_t *mesh_id);
int hpb_mesh_count(const hpb_ctx *ctx);
/* Copies the mip level `level` (RGBA8, tex_w*tex_h*4 bytes) back to the host; for tests. */
int hpb_mesh_get_mip(hpb_ctx *ctx, int32_t mesh_id, int level, uint8_t *out_host, int *w, int *h, int *levels);
/*
 * Closed-surface analysis done at upload.  *sign = -1 / +1 when the mesh (vertices welded by position) is a closed,
 * consistently oriented surface whose front faces project with negative / positive signed area, 0 otherwise.  The
 * rasteriser skips back faces of such meshes: on a closed surface every pixel a back face covers is also covered by
 * a nearer front face, so the two-sided semantics of the reference (panda3d_scene_renderer.py:102, set_two_sided)
 * are unchanged; scenes the near plane cuts open are rendered two-sided.  hpb_mesh_set_cull(ctx, id, 0) forces
 * two-sided rendering of every triangle of the mesh (tests use it to check the equivalence), 1 restores the default.
 */
int hpb_mesh_closed_sign(hpb_ctx *ctx, int32_t mesh_id, int *sign);
int hpb_mesh_set_cull(hpb_ctx *ctx, int32_t mesh_id, int enable);

/*
 * Batched rasteriser: renders b single-object scenes in ONE launch.
 *   mesh_ids_dev [b] int32, TCO_dev [b,16], K_dev [b,9] float32 (device)
 *   ambient_dev  [b,3] summed ambient light colour per scene, or NULL (= 1,1,1)
 *   lights_dev   [b, n_lights, 8] float32 or NULL (n_lights = 0): point / directional lights of each scene,
 *                (type 0 = point | 1 = directional, x, y, z, r, g, b, unused): a point light's POSITION, a directional
 *                light's DIRECTION of travel, in the world frame (= the object frame: the reference renders the object at
 *                the identity and moves the camera, panda3d_batch_renderer.py:144-192).  n_lights <= 8.  Shading is
 *                per-pixel Lambert without attenuation, colour = albedo * min(1, ambient + sum_i c_i * max(0, n . l_i)) with n
 *                the unit eye-space normal: the render_normals=False light rig of pose_rigid.py:105-141,421-422
 *                (1 ambient 0.1 + 6 point lights 0.4 at +-10 bounding radii, panda3d_scene_renderer.py:105-141).
 *   flags        HPB_RENDER_* ; outputs not selected may be NULL
 *   rgb/normals  [b,3,h,w] float32 in {k/255};  depth [b,1,h,w] float32 metres (0 = background);
 *   mask         [b,1,h,w] uint8 0/1.
 *   addressing   scene i writes its planes at  base + (i / views) * bstride + (i % views) * view_stride  (elements).
 *                views = 1 is the plain batched layout.  views = V with view_stride = C_r*h*w lets the V renders of
 *                one hypothesis land side by side in the channels of a [b/V, C_in, h, w] network input
 *                (render_images_multiview's view(bsz, n_views, C, h, w).flatten(1, 2), pose_rigid.py:447-452).
 *                The mask is never view-interleaved: scene i writes at mask_dev + i * mask_bstride.
 * A non-finite TCO or K yields all-zero images for that scene (panda3d_batch_renderer.py:81-111).
 */
int hpb_render(hpb_ctx *ctx, const int32_t *mesh_ids_dev, const float *TCO_dev, const float *K_dev,
               const float *ambient_dev, int b, int h, int w, float z_near, float z_far, uint32_t flags,
               float *rgb_dev, int64_t rgb_bstride, float *normals_dev, int64_t normals_bstride,
               float *depth_dev, int64_t depth_bstride, uint8_t *mask_dev, int64_t mask_bstride,
               int views, int64_t view_stride, const float *lights_dev, int n_lights, void *stream);

/*
 * Perspective crop + resize of the observed frame(s) (crop_inputs).
 *   images_dev [n_im,C,H,W] float32 (C = 3 or 4; channel 3 = depth), im_ids_dev [b] int32 selects the
 *   frame of each row (replaces the reference's images[batch_im_ids] expansion, pose_estimator.py:390)
 *   points_dev [n_obj,n_pts,3] float32 point sets (mesh_db.select(labels).sample_points(2000)),
 *   obj_ids_dev [b] int32 row of points_dev per hypothesis
 *   K_dev [b,9], TCO_dev [b,16], tCR_dev [b,3]
 *   outputs: crops_dev [b,C,h,w] (batch stride crops_bstride elements), K_crop_dev [b,9],
 *            boxes_rend_dev [b,4], boxes_crop_dev [b,4]  (x1,y1,x2,y2)
 *   tap_bits  32 / 16 = precision of the frame samples for THIS launch (see hpb_set_crop_tap_precision), 0 = the
 *             context default.  A launch argument, so concurrent callers with different precisions cannot race.
 */
int hpb_crop(hpb_ctx *ctx, const float *images_dev, int n_im, int C, int H, int W, const int32_t *im_ids_dev,
             const float *points_dev, int n_obj, int n_pts, const int32_t *obj_ids_dev, const float *K_dev,
             const float *TCO_dev, const float *tCR_dev, int b, int h, int w, float lamb, float *crops_dev,
             int64_t crops_bstride, float *K_crop_dev, float *boxes_rend_dev, float *boxes_crop_dev,
             int tap_bits, void *stream);

/*
 * hpb_crop for RGB frames (C = 3) with the crop written as crops_dev [b, h, w] pixels of 4 bfloat16 (r, g, b, 0)
 * (8 bytes per pixel, batch stride crops_bstride pixels): each value is the float32 crop of hpb_crop rounded to nearest
 * even, i.e. exactly what the network input packing would make of it.  Consumer: hpb_render_s2d_bf16(HPB_CROPS_BF16X4).
 */
int hpb_crop_bf16x4(hpb_ctx *ctx, const float *images_dev, int n_im, int H, int W, const int32_t *im_ids_dev,
                    const float *points_dev, int n_obj, int n_pts, const int32_t *obj_ids_dev, const float *K_dev,
                    const float *TCO_dev, const float *tCR_dev, int b, int h, int w, float lamb, void *crops_dev,
                    int64_t crops_bstride, float *K_crop_dev, float *boxes_rend_dev, float *boxes_crop_dev,
                    int tap_bits, void *stream);

/*
 * Precision of the frame samples ("taps") hpb_crop reads when it resamples from its pixel-interleaved copy of an RGB
 * frame (many hypotheses per frame).  32 (default): float32, the reference's arithmetic (cropping.py:155-197, torchvision
 * roi_align on float32 images).  16: the copy holds IEEE fp16 -- 8-byte instead of 16-byte taps, which halves the L1
 * sector look-ups that bound the kernel; the crop of a frame f then equals the 32-bit crop of fp16(f) bit for bit, i.e.
 * an absolute error <= 2.5e-4 for pixel values in [0,1] (BASELINE bar for crops: 1e-3).  Meant for the bf16 network
 * path, whose input is rounded to 8 mantissa bits anyway.  RGB-D frames always use 32.  This sets the context DEFAULT
 * used by launches that pass tap_bits = 0.  Host-side switch, no sync.
 */
int hpb_set_crop_tap_precision(hpb_ctx *ctx, int bits);
/*
 * With 16-bit taps of an RGB frame the crop's source rows are streamed into shared memory by the TMA unit
 * (cp.async.bulk.tensor.3d boxes + mbarrier ring, hpb_crop_tma.cu) instead of being gathered by per-lane loads; results are
 * bit-identical.  Default: DISABLED -- measured on B200 (profiles/r2_tma_kernels.txt) the ring is as fast as the per-lane
 * kernel at 576 rows and 5 % slower at 2304: the crop is bound by its filter arithmetic, not by where the taps come from.
 */
int hpb_set_crop_tma(hpb_ctx *ctx, int enable);

/* Boxes and K_crop only (compute_crops_multiview: return_crops=False, 200 points). H,W = source frame size. */
int hpb_crop_boxes(hpb_ctx *ctx, int H, int W, const float *points_dev, int n_obj, int n_pts,
                   const int32_t *obj_ids_dev, const float *K_dev, const float *TCO_dev, const float *tCR_dev,
                   int b, int h, int w, float lamb, float *K_crop_dev, float *boxes_rend_dev,
                   float *boxes_crop_dev, void *stream);

/*
 * The resampling half of hpb_crop on its own: crops_dev [b,C,h,w] = roi_align of the frames at boxes_crop_dev [b,4]
 * (x1,y1,x2,y2), e.g. the boxes hpb_refiner_prologue produced.  Arguments as in hpb_crop.
 */
int hpb_crop_pixels(hpb_ctx *ctx, const float *images_dev, int n_im, int C, int H, int W, const int32_t *im_ids_dev,
                    const float *boxes_crop_dev, int b, int h, int w, float *crops_dev, int64_t crops_bstride, int tap_bits,
                    void *stream);

/*
 * Fused prologue of one refiner iteration (PosePredictor.forward, pose_rigid.py:570-612), ONE launch for what the reference
 * does in ~60 eager ops plus a per-sample CPU loop: T_norm = normalize_T(TCO); tCR = its translation (reference point =
 * object origin); TCV_O = make_TCO_multiview(T_norm, tCR, mv_type, n_views, remove_tco_rendering); the row's crop geometry
 * (boxes_rend, boxes_crop, K_crop) from the n_pts_crop-point set (crop_inputs :199-277 without the resampling); and
 * KV_crop [b,n_views,9] from the n_pts_mv-point set per view (compute_crops_multiview :279-337) with view 0 = K_crop unless
 * remove_tco_rendering (:610-611).  Bit-identical to hpb_normalize_T -> hpb_multiview -> hpb_crop_boxes -> hpb_crop_boxes.
 * KV_crop_dev may be NULL (single-view models).  points_*_dev [n_obj, n_pts, 3], obj_ids_dev [b].
 */
int hpb_refiner_prologue(hpb_ctx *ctx, const float *TCO_dev, const float *K_dev, const int32_t *obj_ids_dev,
                         const float *points_crop_dev, int n_obj, int n_pts_crop, const float *points_mv_dev, int n_pts_mv, int b,
                         int H, int W, int h, int w, float lamb, int mv_type, int n_views, int remove_tco_rendering,
                         float *T_norm_dev, float *tCR_dev, float *TCV_O_dev, float *K_crop_dev, float *boxes_rend_dev,
                         float *boxes_crop_dev, float *KV_crop_dev, void *stream);

/* T -> normalize_T(T): Gram-Schmidt on columns 0,1 of R, translation kept, last row (0,0,0,1). In place allowed. */
int hpb_normalize_T(hpb_ctx *ctx, const float *T_dev, int b, float *T_out_dev, void *stream);

/* Image-space pose update.  out_dev [b,9] (variants 0,1) or [b,7] (variant 2); tCR_dev [b,3] is only read by
 * variant 0.  TCO_out_dev may alias TCO_dev. */
int hpb_pose_update(hpb_ctx *ctx, const float *TCO_dev, const float *K_crop_dev, const float *out_dev,
                    const float *tCR_dev, int b, int variant, float *TCO_out_dev, void *stream);

/* Initial poses from 2-D boxes.  boxes_dev [b,4]; points_dev [n_obj,n_pts,3] + obj_ids_dev [b] (all mesh points,
 * variants 0,1); K_dev [b,9]; R_dev [b,9] (variant 0); z_mean (variant 2). */
int hpb_tco_init(hpb_ctx *ctx, int variant, const float *boxes_dev, const float *points_dev, int n_obj, int n_pts,
                 const int32_t *obj_ids_dev, const float *K_dev, const float *R_dev, float z_mean, int b,
                 float *TCO_out_dev, void *stream);

/* Extra refiner viewpoints: TCV_O_dev [b,V,16] with V = n_views; view 0 = TCO unless remove_tco_rendering. */
int hpb_multiview(hpb_ctx *ctx, const float *TCO_dev, const float *tCR_dev, int b, int mv_type, int n_views,
                  int remove_tco_rendering, float *TCV_O_dev, void *stream);

/* depth_out = normalize_depth(depth, tCR_z) on `n_planes` [h*w] planes per row; plane p of row n lives at
 * depth_dev + n*bstride + plane_offsets_host[p]*h*w (host array of channel indices). In place. */
int hpb_normalize_depth(hpb_ctx *ctx, float *depth_dev, int64_t bstride, const int32_t *plane_channels_host,
                        int n_planes, const float *tCR_dev, int b, int h, int w, int kind, void *stream);

/*
 * Segmented top-K (filter_top_pose_estimates): keeps, for every group, the K rows with the largest score and
 * returns the surviving row indices in GLOBAL descending-score order (groups interleaved), exactly the order of
 * df.sort_values(field, ascending=False).groupby(cols).head(K).index.  Ties: lowest row index first; NaN last.
 *   scores_dev [n] float32, group_ids_dev [n] int32 in [0,n_groups)
 *   out_idx_dev [min(n, n_groups*K)] int64, out_count_dev [1] int32 (device)
 */
int hpb_topk_segmented(hpb_ctx *ctx, const float *scores_dev, const int32_t *group_ids_dev, int n, int n_groups,
                       int K, int64_t *out_idx_dev, int32_t *out_count_dev, void *stream);

/*
 * Network-input packing: float32 planar x_dev [b,C,h,w] (batch stride x_bstride elements) -> bfloat16 pixel-interleaved
 * out_dev [b,h,w,C_padded] (= a torch channels_last [b,C_padded,h,w] bfloat16 tensor); channels C..C_padded-1 are zero.
 * This is the hand-off from the kernels above to the (unchanged) torch ResNet run in bf16: it replaces the
 * torch.cat((images_crop, renders)) + dtype/layout conversion in front of net_forward (pose_rigid.py:352-374, :629).
 */
int hpb_pack_input_bf16(hpb_ctx *ctx, const float *x_dev, int64_t x_bstride, int b, int C, int h, int w, void *out_dev,
                        int C_padded, void *stream);

/*
 * Space-to-depth variant of hpb_pack_input_bf16 for a 7x7 / stride 2 / pad 3 stem convolution (ResNet conv1,
 * torchvision_resnet.py:211): out_dev [b, H/2+3, W/2+3, C_padded] bfloat16 with
 *   out[n, I, J, (r*2+s)*Cs + c] = xpad[n, c, 2I+r, 2J+s],  xpad = x zero-padded by 3 pixels,  Cs = C_padded / 4,
 * channels c >= C of every sub-pixel block zero, so that conv1 becomes a 4x4 / stride 1 / unpadded convolution over
 * C_padded channels (same sums, 4x deeper reduction per tap; the weight is permuted the same way).  Every sub-pixel's
 * channel block starts on a 16-byte boundary.  H and W even, C_padded >= 4C and a multiple of 32.
 */
int hpb_pack_input_s2d_bf16(hpb_ctx *ctx, const float *x_dev, int64_t x_bstride, int b, int C, int H, int W, void *out_dev,
                            int C_padded, void *stream);

/*
 * hpb_render (rgb + normals, one view per row) fused with hpb_pack_input_s2d_bf16 for the 9-channel coarse / scoring
 * network input of PosePredictor.forward_coarse (pose_rigid.py:708-788: x = cat(images_crop, renders), then the stem
 * of net_forward :352-374): renders the b scenes and writes out_dev [b, h/2+3, w/2+3, C_padded] bfloat16 with
 *   out[n, I, J, (r*2+s)*Cs + c] = xpad[n, c, 2I+r, 2J+s], Cs = C_padded / 4,  xpad = zero-padded-by-3 9-channel input whose
 * channels 0..2 are the crop (the output of hpb_crop / hpb_crop_bf16x4) and channels 3..8 the rendered rgb + normals;
 * channels 9..Cs-1 of every sub-pixel block are zero.  Bit-identical to hpb_crop -> hpb_render into x ->
 * hpb_pack_input_s2d_bf16, without ever materialising the float32 network input.  h and w even, C_padded a multiple of 32,
 * >= 64.
 *   crops_format   HPB_CROPS_F32_PLANAR: crops_dev [b, 3, h, w] float32, batch stride crops_bstride floats;
 *                  HPB_CROPS_BF16X4:     crops_dev [b, h, w] pixels of 4 bfloat16 (r, g, b, 0), batch stride crops_bstride
 *                                        pixels -- hpb_crop_bf16x4's output: half the bytes, one 8-byte load per pixel
 *   pad_prezeroed  non-zero: the caller guarantees that the padding channels 16..Cs-1 of every sub-pixel block of out_dev
 *                  are already zero (a persistent buffer that only this call writes); the kernel then writes only the 32
 *                  bytes per sub-pixel that carry data.  No effect for C_padded = 64 (Cs = 16: nothing to skip).
 */
#define HPB_CROPS_F32_PLANAR 0
#define HPB_CROPS_BF16X4 1
int hpb_render_s2d_bf16(hpb_ctx *ctx, const int32_t *mesh_ids_dev, const float *TCO_dev, const float *K_dev,
                        const float *ambient_dev, int b, int h, int w, float z_near, float z_far, const void *crops_dev,
                        int64_t crops_bstride, int crops_format, void *out_dev, int C_padded, int pad_prezeroed,
                        void *stream);

/*
 * Input stage of the ICP depth refiner (megapose/inference/icp_refiner.py:138-176, 271-289; refiner_utils.py compute_masks):
 * for each of N pose estimates, from the measured depth map of its frame and its rendered depth map (hpb_render with
 * HPB_RENDER_DEPTH at frame resolution), the mask and the two point clouds the reference extracts on the host:
 *   mask    masks_dev[im] when given (uint8 [n_im,H,W]), else the "threshold" mask: measured > 0, rendered > 0,
 *           |measured - rendered| <= depth_delta_thresh
 *   target  (u * d / fx, v * d / fy, d) of the MEASURED depth where 0.2 < d < 5 inside the mask
 *   source  the same of the RENDERED depth at those pixels where rendered > 0
 * with u, v = (x - cx), (y - cy) truncated to int16 (getXYZ keeps them in an int16 table, :107-110).  Points come out in
 * row-major pixel order.  points_*_dev [N, capacity, 3] float32; counts_dev [N, 2] int32 = (n_target, n_source), the true
 * counts even when they exceed `capacity` (points beyond it are dropped).  mask_out_dev [N,H,W] uint8 or NULL;
 * index_tgt_dev / index_src_dev [N, capacity] int32 or NULL: the linear pixel index y * W + x of every emitted point (the
 * caller gathers its host-side normal maps with them).
 * depth_measured_dev [n_im,H,W], depth_rendered_dev [N,H,W], im_ids_dev [N], K_dev [N,9].
 */
int hpb_icp_points(hpb_ctx *ctx, const float *depth_measured_dev, int n_im, const float *depth_rendered_dev,
                   const uint8_t *masks_dev, const int32_t *im_ids_dev, const float *K_dev, int N, int H, int W,
                   float depth_delta_thresh, int64_t capacity, float *points_tgt_dev, float *points_src_dev,
                   int32_t *counts_dev, uint8_t *mask_out_dev, int32_t *index_tgt_dev, int32_t *index_src_dev, void *stream);

/*
 * nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of the ResNet stem (torchvision_resnet.py:215) on bfloat16
 * pixel-interleaved activations: in_dev [b,H,W,C] -> out_dev [b,(H-1)/2+1,(W-1)/2+1,C], C a multiple of 8.
 * Bit-identical to torch.nn.functional.max_pool2d (max is exact; NaN propagates).
 */
int hpb_maxpool3x3s2_bf16_nhwc(hpb_ctx *ctx, const void *in_dev, int b, int H, int W, int C, void *out_dev, void *stream);
/* C = 64 (the stem) is served by a tile kernel whose input box is staged by TMA (cp.async.bulk.tensor.4d + mbarrier);
 * enable = 0 selects the plain 9-loads-per-output kernel for every shape (A/B measurements, tests).  Same results. */
int hpb_set_maxpool_tma(hpb_ctx *ctx, int enable);

/*
 * The ResNet stem `conv1` + `bn1` + `relu` (torchvision_resnet.py:197-214) after batch-norm folding and the space-to-depth
 * rewrite of the 7x7 / stride-2 convolution (megapose/fast_resnet.py s2d_weight): a 4x4 / stride-1 convolution of the
 * space-to-depth input, + bias, ReLU, bfloat16 out -- as a tcgen05 implicit GEMM (TMA-fed, accumulators in tensor memory):
 *   z_dev [b,Hz,Wz,C] bf16 (hpb_render_s2d_bf16 / hpb_pack_input_s2d_bf16), w_dev [O,4,4,C] bf16 (= the channels_last
 *   [O,C,4,4] weight), bias_dev [O] f32  ->  out_dev [b,Hz-3,Wz-3,O] bf16.
 * Served shape: C = 64, O = 64, Hz >= 19, Wz >= 16 (edge tiles are clipped by TMA); anything else returns HPB_ENOTFOUND and
 * the caller keeps its library convolution.  fp32 accumulation like cuDNN's; the summation order differs, so outputs agree to bf16
 * rounding, not bit for bit.
 * k_slice_mask: bit (4 * tap + k), tap = 4 * kh + kw, is set when the weight slice w[:, kh, kw, 16k .. 16k+15] is not all
 * zero; cleared slices are not multiplied (the 7x7 kernel leaves 15 of the 64 slices of its 8x8 space-to-depth footprint
 * empty).  ~0 multiplies everything.
 */
int hpb_stem_conv4x4_relu_bf16_nhwc(hpb_ctx *ctx, const void *z_dev, int b, int Hz, int Wz, int C, const void *w_dev,
                                    const float *bias_dev, int O, uint64_t k_slice_mask, void *out_dev, void *stream);
/*
 * A BasicBlock convolution of ResNet layer1 (torchvision_resnet.py:59-83) after batch-norm folding: 3x3 / stride 1 / pad 1,
 * 64 -> 64 channels, + bias, + residual_dev (the block's identity; NULL for the block's first convolution), ReLU, bfloat16
 * NHWC in and out -- the same tcgen05 construction as the stem kernel (zero padding = TMA's out-of-bounds fill):
 *   x_dev, residual_dev, out_dev [b,H,W,64] bf16; bias_dev [64] f32; w_dev [64 out][9 taps][64] bf16 (= the channels_last
 *   [64,64,3,3] weight) without a residual, [64 out][10 taps][64] WITH one: the nine taps followed by a 64 x 64 identity, through
 *   which the tensor core itself adds the residual tile (bf16 -> float32 and x * 1.0 are exact).
 * out = relu(conv(x, w) + bias + residual), accumulated and summed in float32, rounded once.  Served: C = O = 64, H >= 18,
 * W >= 16; anything else returns HPB_ENOTFOUND and the caller keeps its library convolution.  out_dev may alias residual_dev,
 * not x_dev.
 */
int hpb_conv3x3_bias_relu_bf16_nhwc(hpb_ctx *ctx, const void *x_dev, int b, int H, int W, int C, const void *w_dev,
                                    const float *bias_dev, int O, const void *residual_dev, void *out_dev, void *stream);
/* Variants of hpb_stem_conv4x4_relu_bf16_nhwc (A/B measurements and cross-checks; same results): 1 (default) = one TMA box per tile holding
 * the tile and its halo, the 16 taps are start-address offsets into it, epilogue through a swizzled staging tile + TMA store;
 * 2 = the same with the epilogue storing straight from registers (8 % slower); 0 = one TMA box per tap (16x the L2 -> SM
 * traffic). */
int hpb_set_stem_tc_halo(hpb_ctx *ctx, int mode);

#ifdef __cplusplus
}
#endif
#endif /* HPB200_H */
