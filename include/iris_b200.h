/*
 * iris_b200.h -- C ABI of libiris_b200.so, the B200-native (sm_100a) replacement for the
 * differentiable Monte-Carlo shading hot path of facebookresearch/iris.
 *
 * The reference has no native code and no FFI: its hot path is Python that calls Mitsuba/OptiX
 * (ray casts) and tiny-cuda-nn (BRDF field) and ~300 ATen kernels per estimator call
 * (SURVEY.md section 2a).  Each entry point below therefore replaces a reference PYTHON interface,
 * cited as file:line relative to the reference tree; INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C, no torch types.  All array arguments are DEVICE pointers on the scene's device unless a
 *     parameter is documented as host.  `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream).  Calls are asynchronous on that stream.
 *   - every function returns 0 on success or a negative IrisStatus; iris_last_error() returns a
 *     thread-local message (the reference raises Python exceptions / asserts).
 *   - the library allocates device memory only in iris_scene_create (the BVH); per-call scratch is a
 *     caller-provided workspace whose size iris_*_workspace_bytes reports.
 *   - n == 0 is a successful no-op everywhere (the reference's early returns on empty lane sets,
 *     utils/path_tracing.py:72-73,153-154,241-242,347-348).
 */
#ifndef IRIS_B200_H
#define IRIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum IrisStatus {
    IRIS_OK = 0,
    IRIS_ERR_INVALID = -1,   /* bad argument */
    IRIS_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    IRIS_ERR_NOMEM = -3,
    IRIS_ERR_WORKSPACE = -4  /* workspace too small */
} IrisStatus;

const char *iris_last_error(void);
/* "iris_b200 <version> sm_100a" */
const char *iris_version(void);
/* Binding handshake: what = 0 -> IRIS_ABI_VERSION, 1 -> sizeof(IrisShadeParams), 2 -> sizeof(IrisSampler), 3 -> sizeof(IrisSceneStats),
 * anything else -> -1.  A binding that mirrors the structs by hand (ctypes, cgo) compares these before the first call. */
#define IRIS_ABI_VERSION 3
int64_t iris_abi_info(int what);

/* ------------------------------------------------------------------------------------------------
 * Scene: one triangle mesh + 8-wide compressed BVH resident in HBM.
 * Replaces mitsuba.load_dict({'type':'scene','shape_id':{'type':'obj'|'ply',...}})
 * (train_emitter.py:57-63, bake_shading.py:55-61); prim index = face order of `faces`.
 * verts/faces are HOST pointers.  builder: 1 = on the device (what the Python layer uses): Morton hierarchy, every subtree of <= 4096
 * primitives rebuilt with a binned SAH in shared memory, SAH-optimal collapse to 8-wide -- 25-100 ms for 1M-5M triangles, ray rates of a host SAH tree
 * also on irregular scan-like meshes (iris_set_option "lbvh_sah_treelets" 0 = plain LBVH, "lbvh_sah_top" 1 = SAH over the clusters for
 * the top levels as well); 0 = host binned SAH (0.7-1 s per 1M triangles).  Hits do not depend on the builder.
 * ---------------------------------------------------------------------------------------------- */
typedef struct IrisScene IrisScene;

typedef struct IrisSceneStats {
    int64_t n_tris;
    int64_t n_nodes;        /* 80-byte BVH8 nodes */
    int64_t node_bytes;
    int64_t tri_bytes;      /* 48-byte triangle records */
    float build_ms;
    float sah_cost;
    int32_t max_depth;      /* levels of 8-wide nodes (<= 32, the traversal stack) */
    float bounds_lo[3], bounds_hi[3];
} IrisSceneStats;

int iris_scene_create(const float *verts, int64_t n_verts, const int32_t *faces, int64_t n_faces,
                      int device, int builder, IrisScene **out);
void iris_scene_destroy(IrisScene *scene);
int iris_scene_stats(const IrisScene *scene, IrisSceneStats *out);

/* ------------------------------------------------------------------------------------------------
 * iris_intersect -- closest hit of n rays.  Replaces ray_intersect(scene,xs,ds)
 * (utils/path_tracing.py:17-48): t (n) [inf on miss], prim (n) [-1 on miss], uv (n,2) barycentrics,
 * p (n,3) hit point, nrm (n,3) unit geometric normal flipped toward -d (utils/ops.py:85-96).
 * Any output pointer may be NULL.  Semantics are bit-for-bit those of oracle/intersect.c.
 * ---------------------------------------------------------------------------------------------- */
int iris_intersect(const IrisScene *scene, const float *o, const float *d, int64_t n,
                   float *t, int32_t *prim, float *uv, float *p, float *nrm, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Shading tables (device pointers, owned by the caller).
 * Emitter part: SLFEmitter buffers (model/emitter.py:136-173).  SLF part: VoxelSLF (model/slf.py:30-39)
 * with `inds` narrowed to int32.  Field part: NGPBRDF (model/brdf.py:213-260) = tcnn HashGrid+MLP
 * parameters at fp16 (`grid` N_ENTRIES x 2, `mlp` 9216, layout [W1|W2|W3] row-major [out][in]).
 * ---------------------------------------------------------------------------------------------- */
typedef struct IrisShadeParams {
    /* emitters */
    const int32_t *emitter_of_face;   /* (F)  emitter index or -1            (is_emitter + emitter_idx) */
    const int32_t *face_of_emitter;   /* (K)  triangle_idx                                              */
    const float *emitter_vertices;    /* (K,3,3)                                                        */
    const float *emitter_area;        /* (K)                                                            */
    const float *emitter_pdf;         /* (K)  1/K                                                       */
    const float *emitter_cdf;         /* (K)  fp32 cumsum of emitter_pdf                                */
    const float *radiance;            /* (>=K,3) rows [0,K) are read                     */
    int32_t n_emitters;               /* K */
    int32_t n_faces;                  /* F */
    /* surface light field */
    const int32_t *slf_inds;          /* (H,H,H) [z][y][x], -1 = empty */
    const float *slf_radiance;        /* (n_occ,3) */
    int32_t slf_H;
    float slf_vmin, slf_range;        /* fp32(voxel_min), fp32(voxel_max - voxel_min) */
    /* BRDF field (may be NULL for the bake entry points) */
    const void *grid_f16;             /* __half[N_ENTRIES*2] */
    const void *mlp_f16;              /* __half[9216] */
    float field_vmin, field_range;
} IrisShadeParams;

/* Uniform samples: an explicit buffer U[(lane*stride)+column] (parity mode: identical injected
 * sequences) or, when U == NULL, Philox-4x32-10 keyed by (seed, lane + lane_offset, column/4). */
typedef struct IrisSampler {
    const float *U;
    int32_t stride;
    uint64_t seed;
    uint64_t lane_offset;
} IrisSampler;

/* Fills out[n*dims] with the uniforms the Philox mode would hand to lanes [0,n) (tests use it to run the
 * oracle on the production stream). */
int iris_sampler_fill(uint64_t seed, uint64_t lane_offset, int64_t n, int32_t dims, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Shading-map bake -- the inner loops of bake_shading.py:108-123 (diffuse) and :168-188 (specular):
 * per pixel, spp importance-sampled secondary rays from (position,normal), emitter radiance or SLF at
 * the hit, mean over spp.  mode 0: out0 = Ld (B,3).  mode 1: out0 = Ls0, out1 = Ls1 for `roughness`.
 * wo is ignored for mode 0.  Uses sampler columns 0..1.
 * ---------------------------------------------------------------------------------------------- */
int iris_bake(const IrisScene *scene, const IrisShadeParams *params, int mode, float roughness,
              const float *position, const float *normal, const float *wo, int64_t n_pixels, int32_t spp,
              const IrisSampler *sampler, float *out0, float *out1, void *stream);

/* ------------------------------------------------------------------------------------------------
 * BSDF samplers, one lane per row -- BaseBRDF.sample_diffuse (mode 0, model/brdf.py:78-88), .sample_specular (mode 1, :112-136) and
 * .sample_brdf (mode 2, :177-210) for callers that drive the sampling loop themselves (bake_shading.py:113-114,173-177); the fused
 * estimators inline the same code.  u (n,u_stride): mode 0/1 read columns 0,1 (sample2); mode 2 reads column 0 (sample1) and 1,2.
 * wo, normal (n,3); mat (n,5) = albedo rgb, roughness, metallic (mode 2; mode 1: NULL = the scalar `roughness` level, else mat[:,3]).
 * Out: wi (n,3); pdf (n); w0 (n,3) = ones | F0*fac | brdf/pdf; w1 (n,3) = F1*fac (mode 1).  pdf / w0 / w1 may be NULL.
 * sin / cos / asin / acos are the library's own fixed polynomial definitions (csrc/trig.cuh), so a sampled direction is a
 * reproducible function of its inputs, identical to the one the fused estimators trace.
 * ---------------------------------------------------------------------------------------------- */
int iris_bsdf_sample(int mode, const float *u, int32_t u_stride, const float *wo, const float *normal, const float *mat, float roughness,
                     int64_t n, float *wi, float *pdf, float *w0, float *w1, void *stream);

/* ------------------------------------------------------------------------------------------------
 * BRDF field -- NGPBRDF.forward (model/brdf.py:243-260): mat (n,5) = albedo rgb, roughness, metallic.
 * iris_field_backward accumulates d_params (fp32, [mlp 9216 | grid N_ENTRIES*2], same flat layout as the
 * tcnn parameter vector `mlp.params`) given d_mat (n,5); d_params must be zero-initialised by the caller
 * (or hold a running sum).
 * ---------------------------------------------------------------------------------------------- */
/* Hash-grid level table (32 entries each): returns the total number of grid entries (13 977 056). */
int64_t iris_field_levels(float *scale, uint32_t *res, uint32_t *size, uint32_t *offset);
/* encoded (optional, may be NULL): n x 64 fp16 = 128 bytes per sample, 16-byte aligned.  The forward stores each sample's hash-grid
 * features there; a backward given the same array reads them instead of gathering the grid a second time. */
int iris_field_forward(const IrisShadeParams *params, const float *position, int64_t n, float *mat, void *encoded, void *stream);
int64_t iris_field_backward_workspace_bytes(int64_t n);
int iris_field_backward(const IrisShadeParams *params, const float *position, const float *d_mat, int64_t n,
                        float *d_params, const void *encoded, void *workspace, int64_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * path_tracing_single forward + adjoint (utils/path_tracing.py:320-407), the estimator
 * train_emitter.py:184-189 and initialize.py:175-180 differentiate.
 *   rays: (B,12) = o,d,dxdu,dydv (utils/dataset/synthetic_ldr.py:50-57); L (B,3) = mean over spp.
 *   Sampler columns: 0 du,1 dv,2 e1,3 e2x,4 e2y,5 b1,6 b2x,7 b2y.
 *   record: NULL for inference; otherwise >= iris_single_record_bytes(B,spp) bytes that the adjoint
 *   replays (per sample 96 bytes: emitter rows + coefficients of the three radiance gathers, the 3x3
 *   sparse Jacobian of the sample's radiance wrt albedo/roughness/metallic, the hit point).
 *   encoded: NULL, or >= iris_single_encoded_bytes(B,spp) = 128 bytes per sample, 16-byte aligned.  Given: the forward keeps every
 *   sample's 64 fp16 hash-grid features there and writes the Jacobian words of the record, and iris_single_backward (which must be
 *   handed the same array) can produce d_mat / d_params without gathering the grid again.  NULL: the record is the emitter-gradient-only
 *   form of train_emitter.py -- no Jacobians computed or stored, iris_single_backward with d_params != NULL returns IRIS_ERR_INVALID.
 *   workspace: >= iris_single_workspace_bytes(B,spp): three 16-byte words per sample and the ray queue
 *   of one chunk (the secondary bounce runs as generate -> persistent ray-queue trace -> shade).
 * iris_single_backward: dL (B,3) -> d_radiance (K,3) accumulated; d_mat (B*spp,5) written (feed it,
 * with `x0` = the primary hit positions held in the record, to iris_field_backward via
 * iris_single_backward's d_params argument: when d_params != NULL the field adjoint is run too).
 * ---------------------------------------------------------------------------------------------- */
int64_t iris_single_workspace_bytes(int64_t n_pixels, int32_t spp);
int64_t iris_single_record_bytes(int64_t n_pixels, int32_t spp);
int64_t iris_single_encoded_bytes(int64_t n_pixels, int32_t spp);
int iris_single_forward(const IrisScene *scene, const IrisShadeParams *params, const float *rays,
                        int64_t n_pixels, int32_t spp, const IrisSampler *sampler, float *L,
                        void *record, void *encoded, void *workspace, int64_t workspace_bytes, void *stream);
int iris_single_backward(const IrisShadeParams *params, const float *dL, int64_t n_pixels, int32_t spp,
                         const void *record, const void *encoded, float *d_radiance, float *d_params,
                         void *workspace, int64_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Forward-only estimators that evaluate the BRDF field at every hit (wavefront: trace -> field -> shade per bounce).
 *   iris_path_tracing      path_tracing(..., spp, indir_depth)            utils/path_tracing.py:214-318  (render.py:171-176)
 *   iris_path_tracing_det  path_tracing_det_diff (mode 0) / _det_spec (1) utils/path_tracing.py:50-212   (refine_shading.py:116-122,160-166)
 *                          positions/wis/normals (B,3), prim (B) int32 with -1 = pixel without a hit -> zeros; mode 1 also writes L1
 *   iris_trace_indirect    trace_indirect(position, wo, normal, depth)    utils/path_tracing.py:409-502  -> L (n,3) per lane
 * Sampler columns: path_tracing 0 du,1 dv,2..7 first bounce, then 6 per indirect depth (e1,e2x,e2y,b1,b2x,b2y); det: 0..1 first sample,
 * then 6 per depth; trace_indirect: 6 per depth from column 0.  workspace >= iris_wave_workspace_bytes(lanes), lanes = rows * spp.
 * ---------------------------------------------------------------------------------------------- */
int64_t iris_wave_workspace_bytes(int64_t n_lanes);
int iris_path_tracing(const IrisScene *scene, const IrisShadeParams *params, const float *rays, int64_t n_pixels, int32_t spp,
                      int32_t indir_depth, const IrisSampler *sampler, float *L, void *workspace, int64_t workspace_bytes, void *stream);
int iris_path_tracing_det(const IrisScene *scene, const IrisShadeParams *params, int mode, float roughness_level, const float *positions,
                          const float *wis, const float *normals, const int32_t *prim, int64_t n_pixels, int32_t spp, int32_t indir_depth,
                          const IrisSampler *sampler, float *L0, float *L1, void *workspace, int64_t workspace_bytes, void *stream);
int iris_trace_indirect(const IrisScene *scene, const IrisShadeParams *params, const float *position, const float *wo, const float *normal,
                        int64_t n, int32_t indir_depth, const IrisSampler *sampler, float *L, void *workspace, int64_t workspace_bytes,
                        void *stream);

/* ------------------------------------------------------------------------------------------------
 * Voxel surface-light-field bake on the device -- slf_bake.py:70-145 with model/slf.py:16-61 (the stage that produces the vslf.npz
 * the estimators read).  All arrays are device arrays; `valid` (n bytes, may be NULL = all valid) is ray_intersect's mask.
 *   iris_slf_bounds      running min / max over every coordinate of the valid points (slf_bake.py:73-85).  state: 16 bytes,
 *                        [0],[1] = min, max as floats (valid after every call), [2],[3] internal; reset != 0 on the first call.
 *   iris_slf_mark        occupancy[cell] = 1 for the voxel of every valid point (SpatialHist > 0, :96-113); zero it first.
 *                        cell = ((gz * H) + gy) * H + gx, g = clamp(trunc((x - voxel_min) / voxel_range * H), 0, H-1).
 *   iris_slf_index       inds (H^3 int32) = -1 for empty voxels, else the rank among the occupied ones in raster order -- the order of
 *                        torch.where(mask) (model/slf.py:29-32); n_cells (host) = number of occupied voxels.  Synchronises the stream.
 *   iris_slf_accumulate  sum[inds[cell]] += radiance, count[inds[cell]] += 1 (VoxelSLF.scatter_add, model/slf.py:56-61).
 *   iris_slf_finalize    sum /= max(count, 1) in place (slf_bake.py:138): sum becomes the `radiance` table of IrisShadeParams.
 * ---------------------------------------------------------------------------------------------- */
int iris_slf_bounds(const float *positions, const uint8_t *valid, int64_t n, int32_t reset, float *state, void *stream);
int iris_slf_mark(const float *positions, const uint8_t *valid, int64_t n, float voxel_min, float voxel_range, int32_t H, int32_t *occupancy,
                  void *stream);
int64_t iris_slf_index_workspace_bytes(int32_t H);
int iris_slf_index(const int32_t *occupancy, int32_t H, int32_t *inds, int64_t *n_cells, void *workspace, int64_t workspace_bytes, void *stream);
int iris_slf_accumulate(const float *positions, const uint8_t *valid, const float *radiance, int64_t n, float voxel_min, float voxel_range,
                        int32_t H, const int32_t *inds, float *sum, int32_t *count, void *stream);
int iris_slf_finalize(float *sum, const int32_t *count, int64_t n_cells, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Emitter extraction on the device -- extract_emitter_ldr.py:76-110 (the stage that produces emitter.pth).  Device arrays.
 *   iris_tri_accumulate    tri_sum[prim] += radiance, tri_count[prim] += 1 for every valid ray that hit a triangle (prim from
 *                          iris_intersect; torch_scatter 'sum' in the reference, :85-90); zero both tables first.
 *   iris_emitter_classify  is_emitter[f] = max_c(tri_sum[f][c] / max(tri_count[f], 1)) > threshold           (:92-95)
 *   iris_emitter_geometry  for the K selected faces (emitter_faces, increasing face index = boolean-mask order): the three
 *                          vertices (K,3,3), area = |e1 x e2| / 2 and unit normal e1 x e2 / max(|e1 x e2|, 1e-12)  (:96-100)
 * ---------------------------------------------------------------------------------------------- */
int iris_tri_accumulate(const int32_t *prim, const uint8_t *valid, const float *radiance, int64_t n, int64_t n_faces, float *tri_sum,
                        int32_t *tri_count, void *stream);
int iris_emitter_classify(const float *tri_sum, const int32_t *tri_count, int64_t n_faces, float threshold, uint8_t *is_emitter, void *stream);
int iris_emitter_geometry(const float *verts, const int32_t *faces, const int64_t *emitter_faces, int64_t n_emitters, float *out_vertices,
                          float *out_area, float *out_normal, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Training-step shading from baked maps -- train_brdf_crf.py:193-206 with utils/ops.py:99-119 (lerp_specular):
 *   kd = albedo (1 - metallic), ks = 0.04 (1 - metallic) + albedo metallic,
 *   L = kd * diffuse + ks * lerp(specular0, roughness) + lerp(specular1, roughness)
 * mat (n,5) = (albedo rgb, roughness, metallic) as iris_field_forward writes it; diffuse (n,3); specular0/1 (n,n_levels,3), the
 * maps baked by iris_bake at roughness levels linspace(0.02, 1, n_levels).  Backward: d_mat (n,5) += J^T dL (zero it, or pre-load
 * it with the gradients of the regularisers that act on albedo / roughness / metallic directly, train_brdf_crf.py:213-302); feed
 * d_mat to iris_field_backward.  The maps are data: no gradient.
 * ---------------------------------------------------------------------------------------------- */
int iris_brdf_shading_forward(const float *mat, const float *diffuse, const float *specular0, const float *specular1, int32_t n_levels,
                              int64_t n, float *L, void *stream);
int iris_brdf_shading_backward(const float *mat, const float *diffuse, const float *specular0, const float *specular1, int32_t n_levels,
                               int64_t n, const float *dL, float *d_mat, void *stream);

/* ------------------------------------------------------------------------------------------------
 * EmorCRF -- the camera response applied right after the estimator in every trainer (crf/model_crf.py:68-86,
 * train_emitter.py:191-193): ldr = lerp(crf_c, clip(hdr * exposure, 0, 1)) with crf (3, n_bins) = f0 + weight @ basis sampled on a
 * regular grid over [0,1].  exposure: (n) with exposure_stride 1, or one value with stride 0.  Backward: d_hdr (n,3) written,
 * d_crf (3,n_bins) accumulated (zero it first); d_weight = d_crf @ basis^T is a (3,dim) matmul left to the caller.
 * ---------------------------------------------------------------------------------------------- */
int iris_crf_forward(const float *hdr, const float *exposure, int32_t exposure_stride, const float *crf, int32_t n_bins, int64_t n,
                     float *ldr, void *stream);
int iris_crf_backward(const float *hdr, const float *exposure, int32_t exposure_stride, const float *crf, int32_t n_bins,
                      const float *d_ldr, int64_t n, float *d_hdr, float *d_crf, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Shading-map denoiser -- the post-process bake_shading.py:81,126-131,190-203 applies to every baked map before writing it
 * (`denoiser = mitsuba.OptixDenoiser(img_hw[::-1]); Ld = denoiser(Ld)`).  OptiX's denoiser is a learned network inside an absent
 * third-party library: parity with it is unpinned by nature.  This entry is a deterministic edge-avoiding a-trous wavelet filter
 * (Dammertz et al. 2010; definition in csrc/denoise.cuh, numpy restatement in oracle/denoise.py) over an (height, width, 3) map:
 * `iterations` levels (tap spacing 1, 2, 4, ...), colour edge-stopping sigma_c (halved per level), and -- when the optional guide
 * images are given -- normal (cosine^sigma_n) and position (sigma_x, scene units) edge-stopping from the bake's primary hits.
 * Pixels whose guide normal is (0,0,0) (no primary hit) pass through and are never used as taps.  normal / position may be NULL.
 * `out` must not alias `image`; workspace (iris_denoise_workspace_bytes) is needed for iterations > 1.
 * ---------------------------------------------------------------------------------------------- */
int64_t iris_denoise_workspace_bytes(int32_t height, int32_t width);
int iris_denoise_atrous(const float *image, const float *normal, const float *position, int32_t height, int32_t width, int32_t iterations,
                        float sigma_c, float sigma_n, float sigma_x, float *out, void *workspace, int64_t workspace_bytes, void *stream);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t iris_launch_count(void);

/* Implementation switches (process-wide; every alternative is tested equal to the default).  Defaults first:
 *   "field_forward_impl"      1 tcgen05 + TMEM 128-row tiles | 0 mma.sync warp tiles
 *   "field_backward_impl"     2 dgrad + wgrad fused in one tcgen05 kernel, TMA tile loads / stores, 8 epilogue warps (used when the encoded inputs are
 *                             available) | 1 the first-generation fused kernel (4 warps, per-thread tile copies) | 0 mma.sync dgrad + TF32 wgrad kernels
 *   "single_impl"             1 wavefront bounce (generate -> persistent ray-queue trace -> shade) | 0 one fused kernel
 *   "single_chunk_log2"       23: samples per ray-queue chunk = 2^value (10..30)
 *   "wave_impl"               1 path_tracing / det / indirect bounces through the ray queue | 0 fused bounce kernel
 *   "wave_compact"            1 live-lane lists per bounce (dense ray queue / hit records, kernels exit past the device-side count: the reference's
 *                             lane compaction, utils/path_tracing.py:347-352,492-501) | 0 every bounce kernel over all lanes
 *   "bake_impl"               2 persistent kernel, generator and radiance lookup inside | 0 fused kernel with block-level direction sort | 1 ray queue
 *   "intersect_impl"          0 one ray per lane (best on camera rays) | 1 persistent warps with dynamic ray fetch (best on incoherent rays)
 *   "persist_ctas_per_sm"     8: resident CTAs per SM of the persistent kernels (1..16)
 *   "lbvh_sah_treelets"       1 device builder rebuilds subtrees of <= 4096 primitives with a binned SAH | 0 plain Morton LBVH
 *   "lbvh_sah_top"            0 Morton splits above the treelets | 1 SAH tree over the clusters (slower over whole interior views)
 *   "tc5_ctas_per_sm" 4, "tc5_bwd_ctas_per_sm" 2, "scatter_ctas_per_sm" 0 (uncapped), "field_smem_carveout_pct" 70, "trace_smem_carveout_pct" -1:
 *                             occupancy / L1 tuning of the field and tracing kernels (measurement knobs)
 * Unknown names or out-of-range values return IRIS_ERR_INVALID. */
int iris_set_option(const char *name, int value);

/* Optional per-kernel device timing: when enabled every launch is bracketed by CUDA events on its own stream.
 * iris_profile_read synchronises the pending events and returns the launch count and the summed duration of one
 * kernel class (ids 0..19, names from iris_profile_name; NULL past the end).  Used for bench.py's roofline line. */
int iris_profile_enable(int on);
const char *iris_profile_name(int kernel_id);
int iris_profile_read(int kernel_id, int64_t *launches, double *total_ms, int reset);

#ifdef __cplusplus
}
#endif
#endif /* IRIS_B200_H */
