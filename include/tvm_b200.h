/*
 * tvm_b200.h — C ABI of the B200-native TensoRF-VM ray renderer (libtvm_b200.so).
 *
 * The reference (mbortolon97/IFFNeRF) has no FFI layer: its boundary for this path is
 * two Python callables, TensorBase.forward (models/tensorBase.py:775-917) and
 * OctreeRender_trilinear_fast (renderer.py:12-25).  This header is the boundary a
 * native replacement exposes *underneath* those callables; the Python host in
 * iffnerf_b200/ binds it with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every entry point returns int: 0 = ok, >0 = cudaError_t, <0 = TVM_E_* argument error
 *   - never allocates, never synchronises the device, never throws; the caller owns all
 *     buffers (device pointers unless stated) and passes an explicit cudaStream_t (as void*)
 *   - no global mutable state: re-entrant per stream, one process per GPU
 *   - all floating point is IEEE fp32; `valid_bits` / counts are bit-exact w.r.t. the
 *     reference (sample positions are computed with the reference's un-fused op order)
 */
#ifndef TVM_B200_H
#define TVM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVM_ABI_VERSION 11

/* argument errors (negative so they cannot collide with cudaError_t) */
#define TVM_E_NULL        (-1)   /* required pointer is NULL                      */
#define TVM_E_SHAPE       (-2)   /* unsupported channel count / grid / stride     */
#define TVM_E_WORKSPACE   (-3)   /* workspace too small                           */
#define TVM_E_MODE        (-4)   /* unsupported activation / shading mode         */

/* flags for tvm_render_fwd / tvm_march_bwd / tvm_sample_mask */
#define TVM_F_EARLY_TERM   (1u << 0)  /* eval only: stop a ray once T < early_term_eps                      */
#define TVM_F_MLP_BF16     (1u << 1)  /* shade with the bf16 tensor-core MLP (tolerance 1e-2) instead of fp32 */
#define TVM_F_NO_SHADE     (1u << 2)  /* stop after the march stage: workspace holds ray_feat/acc/depth      */
#define TVM_F_MLP_TC3      (1u << 4)  /* shade on the tensor cores with bf16x3 SPLIT operands (hi.hi + hi.lo + lo.hi,
                                         fp32 accumulate): fp32-equivalent, rgb within ~1e-6 of the FFMA kernel   */
#define TVM_F_SPLIT_APP    (1u << 5)  /* eval: split the march into a sigma-march that emits per-ray appearance sample lists
                                         and an appearance-gather kernel (needs the larger workspace of
                                         tvm_workspace_bytes(..., TVM_F_SPLIT_APP); ignored with per-sample outputs) */
#define TVM_F_ZERO_UNLIT   (1u << 6)  /* with TVM_F_SPLIT_APP: also write zero ray_feat rows for rays without appearance
                                         samples (callers that read the workspace themselves) */
#define TVM_F_MASK_ANYWHERE (1u << 7) /* tvm_sample_mask: the occupancy test alone decides, also for samples outside the
                                         field's aabb (filtering_rays(bbox_only=False), tensorBase.py:728-737) */
#define TVM_F_GATHER_ONLY  (1u << 8)  /* measurement, with TVM_F_SPLIT_APP: run only the appearance-gather stage on the lists
                                         a previous call left in the same workspace */
#define TVM_F_COUNT_FETCH  (1u << 9)  /* measurement, with TVM_F_GATHER_ONLY: app_count[] receives the number of 16-byte texel
                                         fetches the gather issued per ray (16/48-component fields) */
#define TVM_F_BWD_RUNS     (1u << 10) /* tvm_march_bwd, training scatter: quads walk runs of the block's samples and add equal
                                         texel addresses in registers before one red.v4 (-60 % L2 reduction traffic, +71 %
                                         instructions; slower at the measured sizes, opt-in) */
#define TVM_APP_CAP        128        /* entries per ray in the appearance lists; longer rays take the fused kernel */
#define TVM_F_POINT_SAMPLES (1u << 3) /* sampler of sample_point_color (tensorBase.py:623-638): n_samples samples
                                         centred on the ray origin, z_i = stepSize*(i - n_samples/2)          */

/*
 * Field descriptor (POD, passed by pointer from the host; copied into kernel params).
 * Mirrors the scalars TensorBase.__init__/update_stepSize derive
 * (models/tensorBase.py:262-326, :354-375) and the factor shapes of
 * TensorVMSplit.init_one_svd (models/tensoRF.py:160-170).
 *
 * Packed factor buffer (`factors`, fp32, channel-last so one texel is one contiguous
 * run of C floats): for k in 0..2
 *     density plane k : [G[m1]][pitch][n_sigma[k]]   at float offset dplane_off[k]; pitch = G[m0] | 1 (rows padded
 *                       to an odd texel count so vertically adjacent texels use different L1 bank halves)
 *     density line  k : [G[v]][n_sigma[k]]           at dline_off[k]
 *     app plane k     : [G[m1]][pitch][n_app[k]]     at aplane_off[k]
 *     app line  k     : [G[v]][n_app[k]]             at aline_off[k]
 * with matMode (m0,m1) = (0,1),(0,2),(1,2) and vecMode v = 2,1,0 (tensorBase.py:311-312).
 * Channel counts must be multiples of 4, n_sigma[k] <= 16, n_app[k] <= 48.
 */
typedef struct tvm_field_desc {
    float    aabb[6];            /* lo.xyz, hi.xyz                      (self.aabb)                        */
    float    inv_aabb[3];        /* 2 / aabbSize                        (self.invaabbSize, :358)           */
    int32_t  grid[3];            /* gridSize x,y,z                      (self.gridSize)                    */
    float    step_size;          /* self.stepSize.item()                (:366)                             */
    float    near_t, far_t;      /* (float)near_far[0], [1]             (:498, clamp :502)                 */
    float    density_shift;      /* (:752)                                                                 */
    float    distance_scale;     /* (:849)                                                                 */
    float    weight_thres;       /* rayMarch_weight_thres               (:851)                             */
    float    early_term_eps;     /* T threshold for TVM_F_EARLY_TERM (0 disables)                          */
    int32_t  act;                /* 0 = softplus(beta 1, threshold 20), 1 = relu   (:750-754)              */
    int32_t  n_sigma[3];         /* density components per plane                                           */
    int32_t  n_app[3];           /* appearance components per plane                                        */
    int32_t  app_dim;            /* basis_mat rows (27)                                                    */
    int32_t  fea_pe, view_pe;    /* MLPRender_Fea frequencies (:165-183)                                   */
    int32_t  feature_c;          /* hidden width (128)                                                     */
    int64_t  dplane_off[3], dline_off[3], aplane_off[3], aline_off[3];   /* float offsets into `factors`   */
    int64_t  n_factor_floats;    /* total floats in `factors`                                              */
    /* occupancy (AlphaGridMask, tensorBase.py:50-83); occ_cells == NULL -> no mask                        */
    const uint8_t* occ_cells;    /* [Dz][Dy][Dx] corner codes from tvm_pack_occupancy                      */
    int32_t  occ_dims[3];        /* Dx, Dy, Dz                                                             */
    float    occ_lo[3];          /* alphaMask.aabb[0]                                                      */
    float    occ_inv[3];         /* alphaMask.invgridSize = 1/aabbSize*2 (:59)                             */
    const uint8_t* occ_coarse;   /* [cz][cy][cx] 16^3 super-cells: any occupied corner inside (same buffer)  */
    int32_t  occ_cdims[3];       /* cx, cy, cz = ceil(D/16)                                                */
    /* device pointers to parameters                                                                       */
    const float* factors;        /* packed factors (layout above)                                          */
    const float* basis;          /* basis_mat.weight [app_dim][sum(n_app)] row-major (torch layout)        */
    const float* mlp;            /* packed MLP from tvm_pack_mlp                                           */
    const void*  mlp_tc;         /* bf16 tensor-core weight images from tvm_pack_mlp_tc (NULL unless
                                    TVM_F_MLP_BF16 is used)                                                */
    const void*  mlp_tc3;        /* split (hi|lo) bf16 images from tvm_pack_mlp_tc3 (NULL unless TVM_F_MLP_TC3) */
} tvm_field_desc;

int  tvm_abi_version(void);
/* number of kernels this library has launched in the process so far (all entry points) */
unsigned long long tvm_launch_count(void);
/* last error string for a returned code (static storage) */
const char* tvm_error_string(int code);

/* ---- layout conversion (reference layout <-> kernel layout) -------------------------------------- */

/* Reference NCHW parameters -> packed channel-last buffer.  planes/lines: 6 device pointers each,
 * order density[0..2], app[0..2]; shapes [1,C,G[m1],G[m0]] and [1,C,G[v],1] (tensoRF.py:165-168). */
int tvm_pack_factors(const tvm_field_desc* desc, const float* const planes[6], const float* const lines[6],
                     float* packed, void* stream);
/* Inverse, for gradients: packed grad buffer -> 12 NCHW tensors; accumulate != 0 adds into them
 * (so torch-side regularisers, train.py:299-325, keep adding into the same .grad). */
int tvm_unpack_factor_grads(const tvm_field_desc* desc, const float* packed_grad, float* const planes[6],
                            float* const lines[6], int accumulate, void* stream);
/* The same with the values multiplied by `scale` on the way out (1/world after a data-parallel SUM all-reduce) and,
 * when rezero != 0, the packed buffer cleared behind the read so that it can be scattered into again without a
 * separate memset (a persistent gradient workspace). */
int tvm_unpack_factor_grads_scaled(const tvm_field_desc* desc, float* packed_grad, float* const planes[6],
                                   float* const lines[6], int accumulate, float scale, int rezero, void* stream);
/* alphaMask volume [Dz][Dy][Dx] fp32 (>0 = occupied) -> `cells`: per-cell 8-corner codes [Dz][Dy][Dx] (bytes)
 * followed, at byte offset tvm_occupancy_coarse_offset(), by the 16^3 super-cell summary [cz][cy][cx] used to
 * skip empty space.  `cells` must hold tvm_occupancy_bytes() bytes. */
size_t tvm_occupancy_bytes(int dx, int dy, int dz);
size_t tvm_occupancy_coarse_offset(int dx, int dy, int dz);
int tvm_pack_occupancy(const float* volume, int dx, int dy, int dz, uint8_t* cells, void* stream);
/* MLPRender_Fea weights (torch layout: w1 [C,in], w2 [C,C], w3 [3,C]) -> kernel layout. */
size_t tvm_mlp_pack_floats(const tvm_field_desc* desc);
int tvm_pack_mlp(const tvm_field_desc* desc, const float* w1, const float* b1, const float* w2, const float* b2,
                 const float* w3, const float* b3, float* packed, void* stream);

/* bf16 operand images (K-major canonical core-matrix layout) of basis_mat and the three MLP weights for the
 * tcgen05 shade kernel selected by TVM_F_MLP_BF16; biases are read from the fp32 pack (desc->mlp). */
size_t tvm_mlp_tc_pack_bytes(const tvm_field_desc* desc);
int tvm_pack_mlp_tc(const tvm_field_desc* desc, const float* basis, const float* w1, const float* w2, const float* w3,
                    void* packed, void* stream);

/* the same four operands as 2-term bf16 splits (hi|lo image pairs) for TVM_F_MLP_TC3 */
int tvm_mlp_tc3_supported(const tvm_field_desc* desc);   /* 1 if the head's shape fits the split kernel's smem */
size_t tvm_mlp_tc3_pack_bytes(const tvm_field_desc* desc);
int tvm_pack_mlp_tc3(const tvm_field_desc* desc, const float* basis, const float* w1, const float* w2, const float* w3,
                     void* packed, void* stream);

/* Output placement for ray-sharded renders: ray i of the call is written at offset dst_index[i] (i when
 * dst_index is NULL) of EVERY destination image — rgb[k] [N][3] and depth[k] [N] for k < n_dst, typically the image
 * buffers of all ranks mapped through NVLink peer memory (CUDA IPC / symmetric memory), so the shading epilogue is the
 * all-gather.  All pointers are device-accessible from the launching GPU. */
#define TVM_MAX_PEERS 8
typedef struct tvm_scatter_out {
    const int64_t* dst_index;          /* device [n_rays] destination ray offsets, or NULL                        */
    int32_t        n_dst;              /* 1 .. TVM_MAX_PEERS                                                     */
    float*         rgb[TVM_MAX_PEERS];
    float*         depth[TVM_MAX_PEERS];
} tvm_scatter_out;

/* ---- the hot path ---------------------------------------------------------------------------------- */

/* sample_ray + aabb clip + alphaMask test (tensorBase.py:494-536, :832-837), no factor access.
 * n_samples <= 4096.  rays: [n_rays][ray_stride] fp32 (cols 0-2 origin, 3-5 direction); jitter: NULL (eval) or [n_rays]
 * (train, one U[0,1) per ray, :507-509).  valid_bits: [n_rays][ceil(n_samples/32)] little-endian bit i%32
 * of word i/32 = ray_valid[i] (nullable); counts: [n_rays] = popcount (nullable). */
int tvm_sample_mask(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride, int n_samples,
                    const float* jitter, uint32_t flags, uint32_t* valid_bits, int32_t* counts, void* stream);

/* bytes of scratch tvm_render_fwd needs for n_rays (ray_feat [n][sum(n_app)], acc, depth, counters). */
int tvm_workspace_bytes(const tvm_field_desc* desc, int64_t n_rays, uint32_t flags, size_t* out);

/* TensorBase.forward (tensorBase.py:775-917) for one batch of rays, sample_ray branch.
 * bg: DEVICE pointer to the 3 background floats (train.py:267-271 draws it on the device).
 * Outputs (device, caller-allocated): rgb [n][3], depth [n], acc [n]; optional alpha/z_vals/dists
 * [n][n_samples] (train outputs; any may be NULL), optional valid_bits/valid_count/app_count as above.
 * `ws` keeps the per-ray accumulations the backward (tvm_shade_bwd + tvm_march_bwd) needs. */
int tvm_render_fwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride, int n_samples,
                   const float* jitter, const float* bg /* device [3] */, uint32_t flags,
                   float* rgb, float* depth, float* acc,
                   float* alpha, float* z_vals, float* dists,
                   uint32_t* valid_bits, int32_t* valid_count, int32_t* app_count,
                   void* ws, size_t ws_bytes, void* stream);

/* Backward of the march stage: given d(ray_feat) [n][sum(n_app)], d(acc) [n] and d(alpha) [n][n_samples]
 * (any NULL = zero) it re-marches the rays and scatters into g_factors (packed layout, pre-zeroed or
 * accumulating) and, when g_rays != NULL, writes d(rays) [n][6] (pose mode).  With TVM_F_EARLY_TERM (and
 * d_alpha == NULL) the re-march stops at T < early_term_eps like the forward it pairs with.  ws is the forward workspace
 * (tvm_render_fwd with the same rays / n_samples / jitter; replaces autograd through grid_sampler_2d_backward,
 * cumprod, softplus ... driven by train.py:338 and inerf/estimate_pose_inerf.py:178). */
int tvm_march_bwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride, int n_samples,
                  const float* jitter, uint32_t flags, const float* d_ray_feat, const float* d_acc,
                  const float* d_alpha, float* g_factors, float* g_rays, const void* ws, size_t ws_bytes,
                  void* stream);

/* Shade stage alone (basis_mat + MLPRender_Fea + background blend + depth tail, tensorBase.py:886-908),
 * reading ray_feat/acc/depth partials from ws. */
int tvm_shade_fwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                  const float* bg /* device [3] */,
                  uint32_t flags, float* rgb, float* depth, float* acc, const void* ws, size_t ws_bytes,
                  void* stream);

/* Backward of the shade stage (replaces autograd through basis_mat, MLPRender_Fea and the blend/clamp,
 * tensorBase.py:886-904).  d_rgb [n][3] (+ optional upstream d_acc_in [n] on the acc_map output) ->
 * d_ray_feat [n][sum(n_app)] and d_acc [n] (the inputs of tvm_march_bwd), optional d_view [n][3] (pose mode) and,
 * when non-NULL, ACCUMULATED parameter gradients g_basis [app_dim][sum(n_app)] (torch layout) and g_mlp
 * (tvm_mlp_grad_floats() floats, packed layout of tvm_pack_mlp; caller zeroes them). */
int tvm_shade_bwd(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                  const float* bg /* device [3] */, const float* d_rgb, const float* d_acc_in, float* d_ray_feat,
                  float* d_acc, float* g_basis, float* g_mlp, float* d_view, const void* ws, size_t ws_bytes,
                  void* stream);
size_t tvm_mlp_grad_floats(const tvm_field_desc* desc);
/* packed MLP gradient -> torch-layout tensors (w1 [C,in], b1, w2 [C,C], b2, w3 [3,C], b3); accumulate != 0 adds */
int tvm_unpack_mlp_grads(const tvm_field_desc* desc, const float* packed_grad, float* w1, float* b1, float* w2,
                         float* b2, float* w3, float* b3, int accumulate, void* stream);

/* Point queries of the density field (no rays).  mode 0: `points` are normalised coordinates, out = raw sigma feature
 * (TensorVMSplit.compute_densityfeature, tensoRF.py:216-235).  mode 1: `points` are world coordinates, out =
 * 1 - exp(-sigma*length) gated by the alphaMask (TensorBase.compute_alpha, tensorBase.py:756-773). */
int tvm_point_density(const tvm_field_desc* desc, const float* points /* [n][3] */, int64_t n_points, int mode,
                      float length, float* out /* [n] */, void* stream);

/* `Ref` shading head (models/ref.py:48-155; what configs/lego.txt:25 / truck.txt:26 train with), eval forward.
 * `params` is ONE device float buffer, rows of the in_c-input matrices padded to in4 = round_up(in_c, 4) floats:
 *   [small_w]  10 x in4 : normal_mlp.0 (3 rows), tint_color_mlp.0 (3), roughness_mlp.0 (1), diffuse_color_mlp.0 (3)
 *   [bott_w]   feature_c x in4 : bottleneck_mlp
 *   [small_b]  10 biases in the same order (+2 pad)      [bott_b] feature_c biases
 *   [spec_w]   3 x (feature_c + 2 n_pairs + 1) : specular_mlp.0 weight, torch layout   [spec_b] 3 (+1 pad)
 *   [ide_mat]  (l_max + 1) x n_pairs : dir_enc_fn.mat (ref_utils.py:85-98)
 * tvm_ref_head_layout() returns the seven float offsets (and in4) so the host packs without duplicating the rule. */
typedef struct tvm_ref_head {
    const float* params;
    int32_t in_c;                /* app_dim (27)                                                        */
    int32_t feature_c;           /* bottleneck width (128)                                              */
    int32_t n_pairs;             /* (m, l) pairs of the integrated directional encoding (19 at deg 4)   */
    int32_t l_max;               /* 2^(deg_view-1)                                                      */
    int32_t m[32], l[32];        /* dir_enc_fn.ml_array rows                                            */
    float   rgb_premultiplier, rgb_bias, rgb_padding;     /* ref.py:60-62, :84-92                       */
    float   diffuse_shift;       /* -log(3)  (ref.py:69)                                                */
    float   rough_shift;         /* -1       (ref.py:75)                                                */
} tvm_ref_head;
size_t tvm_ref_head_floats(const tvm_ref_head* head);
int tvm_ref_head_layout(const tvm_ref_head* head, int32_t offs[8]);
/* tvm_shade_fwd / tvm_shade_ref_fwd with the results placed through `out` (see tvm_scatter_out) instead of into local
 * rgb / depth arrays; acc (local, per call ray) stays optional. */
int tvm_shade_fwd_scatter(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride,
                          const float* bg /* device [3] */, uint32_t flags, const tvm_scatter_out* out, float* acc,
                          const void* ws, size_t ws_bytes, void* stream);
int tvm_shade_ref_fwd_scatter(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays, int64_t n_rays,
                              int ray_stride, const float* bg /* device [3] */, const tvm_scatter_out* out, float* acc,
                              const void* ws, size_t ws_bytes, void* stream);

/* Per-ray tail with this head (tensorBase.py:886-908), reading the march-stage partials from ws like tvm_shade_fwd. */
int tvm_shade_ref_fwd(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays, int64_t n_rays,
                      int ray_stride, const float* bg /* device [3] */, float* rgb, float* depth, float* acc,
                      const void* ws, size_t ws_bytes, void* stream);

/* Backward of that tail (training with the Ref head): d_rgb [n][3] (+ optional upstream d_acc_in [n]) -> d_ray_feat
 * [n][sum(n_app)], d_acc [n], optional d_view [n][3]; parameter gradients are ACCUMULATED into g_basis [in_c][sum(n_app)]
 * and g_params (packed layout of tvm_ref_head.params; the ide_mat section is not touched).  ray_feat / acc / app_count
 * are the march-stage outputs (tvm_render_fwd with TVM_F_NO_SHADE). */
int tvm_shade_ref_bwd(const tvm_field_desc* desc, const tvm_ref_head* head, const float* rays, int64_t n_rays,
                      int ray_stride, const float* bg /* device [3] */, const float* ray_feat, const float* acc,
                      const int32_t* app_count, const float* d_rgb, const float* d_acc_in, float* d_ray_feat,
                      float* d_acc, float* d_view, float* g_basis, float* g_params, void* stream);

/* TensorVMSplit.compute_appfeature (tensoRF.py:237-256): appearance feature basis_mat(plane (x) line) at NORMALISED
 * points [n][3] (zero-padded taps) -> out [n][app_dim].  Used by pose_estimation/sampling.py:535-541 (normals of the
 * Ref head at surface samples). */
int tvm_point_appfeature(const tvm_field_desc* desc, const float* points /* [n][3] */, int64_t n_points,
                         float* out /* [n][app_dim] */, void* stream);

/* Fused ray generation for the pixels a step renders, with the pose gradient.  Replaces
 * get_ray_directions_Ks + get_rays (ray_utils.py:28-100) + the pixel indexing / F.normalize / cat of
 * inerf/estimate_pose_inerf.py:96-99,149-164 (flags = 3) and of dataLoader/blender.py:69-72,105-114 (flags = 1).
 *   c2w        device [P][pose_stride] row-major camera-to-world matrices (rows 0..2 of a [3|4][4]); pose_stride 12 or 16
 *   kinv_host  HOST [9] = inverse intrinsics, row-major (torch.inverse(K), ray_utils.py:50)
 *   pixels     device [n][2] int32 (x, y), or NULL: ray i is pixel (i % width, i / width) of a full image
 *   pose_index device [n] int32 pose of each ray, or NULL (pose 0)
 *   flags      bit 0: normalise the camera-space direction before rotating; bit 1: F.normalize the world direction
 *   rays       device [n][7] = (origin, direction, radii)  (radii as ray_utils.py:90-98)
 * tvm_pixel_rays_bwd ACCUMULATES d(rays) [n][g_stride >= 6] (column 6 = d(radii) when g_stride > 6) into
 * g_c2w [P][3][4] (the autograd edge pose -> rays of estimate_pose_inerf.py:149-178). */
int tvm_pixel_rays_fwd(const float* c2w, int pose_stride, const float* kinv_host, const int32_t* pixels,
                       const int32_t* pose_index, int width, int64_t n, int flags, float* rays, void* stream);
int tvm_pixel_rays_bwd(const float* c2w, int pose_stride, const float* kinv_host, const int32_t* pixels,
                       const int32_t* pose_index, int width, int64_t n, int flags, const float* g_rays, int g_stride,
                       float* g_c2w, void* stream);

/* ---- the one exchange step (csrc/collective.cu) --------------------------------------- */

/* SUM all-reduce of a flat fp32 buffer that every rank holds in NVLink peer memory: peer_bufs[k] (HOST array of n_ranks
 * device pointers) is rank k's buffer as addressable from this GPU, multicast the NVSwitch multicast address of the same
 * buffer or NULL.  Rank `rank` reduces its 1/n_ranks slice over all ranks and stores the result into every rank's
 * buffer (multimem.ld_reduce / multimem.st when multicast != NULL, else peer loads / stores).  n_floats % 4 == 0.
 * The caller brackets the call with a device-side barrier across the ranks (all producers done / all slices written). */
int tvm_allreduce_sum_peer(void* const* peer_bufs, int n_ranks, int rank, int64_t n_floats, void* multicast,
                           void* stream);

/* ---- grid maintenance (csrc/gridops.cu) ------------------------------------------- */

/* getDenseAlpha + updateAlphaMask (tensorBase.py:643-696) in one call: alpha = compute_alpha(dense_xyz, length) on the
 * gx x gy x gz lattice dense_xyz = aabb0 (1 - s) + aabb1 s (lin_* = DEVICE copies of torch.linspace(0, 1, g), so the
 * lattice is bit-identical to the reference's), clamp(0,1), 3x3x3 max-pool, threshold (>= thres -> 1 else 0) ->
 * volume [gz][gy][gx] fp32 {0,1} (the layout AlphaGridMask keeps, :670-681), and box_out[7] = min xyz | max xyz of the
 * occupied lattice points (:686-691) | number of occupied voxels.  The query is gated by the descriptor's CURRENT
 * occupancy (:757-761).  ws: tvm_dense_alpha_workspace_bytes(gx, gy, gz). */
size_t tvm_dense_alpha_workspace_bytes(int gx, int gy, int gz);
int tvm_dense_alpha_mask(const tvm_field_desc* desc, const float* lin_x, const float* lin_y, const float* lin_z,
                         int gx, int gy, int gz, float length, float thres, float* volume, float* box_out /* [7] */,
                         void* ws, size_t ws_bytes, void* stream);

/* One factor of up_sampling_VM / shrink (tensoRF.py:258-316), reference layout in and out: src [channels][h][w] ->
 * dst [channels][h2][w2].  mode 0: F.interpolate(mode="bilinear", align_corners=True); mode 1: the crop
 * dst[c][y][x] = src[c][y + y_off][x + x_off]. */
int tvm_resize_factor(const float* src, int channels, int h, int w, float* dst, int h2, int w2, int mode, int y_off,
                      int x_off, void* stream);

/* filtering_rays(bbox_only=True) (tensorBase.py:716-726): out[i] = 1 iff the slab interval of ray i against the aabb is
 * non-empty (t_max > t_min, no near/far clamp). */
int tvm_rays_hit_box(const tvm_field_desc* desc, const float* rays, int64_t n_rays, int ray_stride, uint8_t* out,
                     void* stream);

/* workspace layout helpers (byte offsets inside ws for n_rays) so the host can view the march outputs */
int tvm_workspace_layout(const tvm_field_desc* desc, int64_t n_rays, size_t* ray_feat_off, size_t* acc_off,
                         size_t* depth_off, size_t* sigma_count_off, size_t* app_count_off, size_t* occ_count_off);

#ifdef __cplusplus
}
#endif
#endif /* TVM_B200_H */
