// iris_lib.cu -- the C ABI of libiris_b200.so (include/iris_b200.h): argument checking, launches, error strings.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/iris_b200.h"
#include "bvh8.h"
#include "bvh_device.cuh"
#include "field.cuh"
#include "field_tc5.cuh"
#include "field_bwd_tc5.cuh"
#include "field_bwd_tc5v2.cuh"
#include "kernels.cuh"
#include "wavefront.cuh"
#include "crf.cuh"
#include "shade_maps.cuh"
#include "slf_bake.cuh"
#include "emitter_extract.cuh"
#include "bsdf_api.cuh"
#include "denoise.cuh"

struct IrisScene {
    int device = 0;
    float4 *nodes = nullptr;
    float4 *tris = nullptr;
    IrisSceneStats stats{};
};

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(x)                                                                                         \
    do {                                                                                                    \
        cudaError_t e__ = (x);                                                                              \
        if (e__ != cudaSuccess) return fail(IRIS_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define LAUNCHED()                                                                                                   \
    do {                                                                                                             \
        g_launches.fetch_add(1);                                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                                        \
        if (e__ != cudaSuccess) return fail(IRIS_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
    } while (0)

// ---- optional per-kernel timing (CUDA events on the launching stream), for bench.py's roofline line
enum KernelId { K_INTERSECT = 0, K_BAKE_DIFFUSE, K_BAKE_SPECULAR, K_PRIMARY, K_FIELD_FORWARD, K_BOUNCE_SINGLE, K_SINGLE_BACKWARD, K_FIELD_BACKWARD, K_FIELD_WGRAD, K_FIELD_SCATTER, K_WAVE_INIT, K_WAVE_A, K_WAVE_B, K_WAVE_FINISH, K_SINGLE_GEN, K_TRACE_QUEUE, K_SINGLE_SHADE, K_BAKE_GEN, K_BAKE_SHADE, K_FIELD_BACKWARD_TC5, K_COUNT };
static const char *g_kernel_names[K_COUNT] = {"k_intersect", "k_bake<0>", "k_bake<1>", "k_primary", "k_field_forward", "k_bounce_single", "k_single_backward", "k_field_backward_dgrad", "k_field_backward_wgrad", "k_field_backward_scatter", "k_wave_init", "k_wave_bounce_a", "k_wave_bounce_b", "k_wave_finish", "k_single_gen", "k_trace_queue", "k_single_shade", "k_bake_gen", "k_bake_shade", "k_field_backward_tc5"};
struct ProfSpan { int id; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfSpan> g_spans;
static std::vector<cudaEvent_t> g_free_events;
static double g_prof_ms[K_COUNT] = {0};
static int64_t g_prof_n[K_COUNT] = {0};
static cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!g_free_events.empty()) { e = g_free_events.back(); g_free_events.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    int id; cudaStream_t st; cudaEvent_t a{}, b{}; bool on;
    ProfScope(int id_, cudaStream_t s) : id(id_), st(s), on(g_prof_on) { if (on) { a = prof_event(); b = prof_event(); cudaEventRecord(a, st); } }
    ~ProfScope() { if (on) { cudaEventRecord(b, st); g_spans.push_back({id, a, b}); } }
};

static inline unsigned blocks_for(int64_t n) { return (unsigned)((n + IRIS_BLOCK - 1) / IRIS_BLOCK); }
static inline SceneView view_of(const IrisScene *s) {
    SceneView v;
    v.nodes = s->nodes;
    v.tris = s->tris;
    v.n_tris = s->stats.n_tris;
    return v;
}

static int g_sm_count = 0;
static int g_tc5_ctas = 4;      // tcgen05 kernel: CTAs per SM (34.9 KB smem, 64 TMEM columns each)
static int g_single_impl = 1;      // 1: wavefront bounce (k_single_gen -> k_trace_queue -> k_single_shade), 0: fused k_bounce_single
static int64_t g_single_chunk = 8 << 20;   // samples per wavefront chunk
static int g_field_bwd_impl = 2;   // 1: fused tcgen05 dgrad + wgrad kernel (field_bwd_tc5.cuh; used whenever the encoded inputs are available); 0: dgrad (mma.sync) + wgrad (TF32 split-K) kernels
static int g_bake_impl = 2;        // 2: persistent warps with the generator / radiance lookup in the kernel (k_bake_persistent, default: +20-44 % over 0
                                   //    except on mirror-like lobes), 0: fused k_bake with block-level direction sort, 1: through the ray queue
static int g_wave_impl = 1;        // 1: wavefront bounces through the ray queue, 0: fused k_wave_bounce_a
static int g_wave_compact = 1;     // 1: live-lane lists (dense queues, dead lanes cost nothing), 0: every kernel over all lanes (A/B)
static int g_intersect_impl = 0;   // 1: persistent warps with dynamic ray fetch (k_intersect_persistent)
static int g_sah_top = 0;          // device builder: 1 = SAH tree over the clusters for the levels above the treelets (host, tiny); 0 = Morton splits
                                   // (default: over whole interior views the Morton top traces 3-7 % faster than the SAH top, profiles/r2t_*)
static int g_sah_treelets = 1;     // device builder: rebuild the lower levels with a binned SAH (bvh_device.cuh); 0 = plain LBVH
static int g_tc5_bwd_ctas = 2;     // fused field adjoint: CTAs per SM (98 KB of shared memory each)
static int g_scatter_ctas = 0;     // > 0: cap the grid of the grid-gradient scatter at this many CTAs per SM (it strides over the samples)
static int g_persist_ctas = 8;     // resident CTAs per SM for the persistent grid
static int g_field_impl = 1;   // 1: tcgen05 / TMEM 128-row tiles (default), 0: mma.sync warp tiles (kept for A/B measurements)
static bool g_field_ready[64] = {false};
static int ensure_device_setup(int device) {
    if (device < 0 || device >= 64) return fail(IRIS_ERR_INVALID, "device index out of range");
    if (g_field_ready[device]) return IRIS_OK;
    FieldLevel lv[FIELD_LEVELS];
    field_level_table(lv);
    for (int l = 0; l < FIELD_LEVELS; ++l) {      // the kernels hard-code the geometry of the dense levels and 2^19 for the hashed ones
        const bool dense = l < FIELD_DENSE_LEVELS;
        if ((lv[l].dense != 0) != dense || (dense && (lv[l].res != field_dense_res(l) || lv[l].size != field_dense_size(l))) ||
            (!dense && lv[l].size != (1u << 19)))
            return fail(IRIS_ERR_INVALID, "hash-grid level table does not match the compiled-in configuration");
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_levels, lv, sizeof(lv)));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    g_sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaFuncSetAttribute(k_single_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * IRIS_BWD_KMAX * IRIS_BLOCK * 4));
    g_field_ready[device] = true;
    return IRIS_OK;
}

// ---- TMA tensor map over a row-major [rows][64] fp16 array with 8-column x 128-row boxes (field_bwd_tc5v2.cuh).  The encoder is a
//      driver-API entry point, fetched through the runtime so that the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_row_tile_map(CUtensorMap *out, const void *base, int64_t rows) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(IRIS_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {8, 128};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(IRIS_ERR_CUDA, "cuTensorMapEncodeTiled failed (status " + std::to_string((int)r) + ")");
    return IRIS_OK;
}

// ---- builder = 1: everything on the device (bvh_device.cuh)
#define CUDA_OK(x)                                              \
    do {                                                        \
        e = (x);                                                \
        if (e != cudaSuccess) goto done;                        \
    } while (0)
static cudaError_t device_bvh_build(const float *verts, int64_t n_verts, const int32_t *faces, int64_t n_faces, IrisScene *s) {
    cudaError_t e = cudaSuccess;
    const int n = (int)n_faces;
    float *d_verts = nullptr, *d_bounds = nullptr;
    int32_t *d_faces = nullptr;
    TriRecord *d_recs = nullptr, *d_tris = nullptr;
    DBox *d_tbox = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_idx = nullptr, *d_idx2 = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    LbvhNodes N{};
    int *d_wide_bin = nullptr, *d_wide_depth = nullptr, *d_counters = nullptr, *d_list = nullptr, *d_nlist = nullptr, *d_rec = nullptr;
    int2 *d_cl = nullptr;
    int3 *d_perm = nullptr, *d_roots = nullptr;
    float *d_cbox = nullptr;
    Bvh8Node *d_wide = nullptr;
    int counters[3] = {1, 0, 1};
    const float binit[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float hb[6];
    const unsigned gb = (unsigned)((n + 255) / 256);
    CUDA_OK(cudaMalloc(&d_verts, sizeof(float) * 3 * (size_t)std::max<int64_t>(n_verts, 1)));
    CUDA_OK(cudaMalloc(&d_faces, sizeof(int32_t) * 3 * (size_t)n));
    CUDA_OK(cudaMemcpy(d_verts, verts, sizeof(float) * 3 * (size_t)n_verts, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_faces, faces, sizeof(int32_t) * 3 * (size_t)n, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&d_recs, sizeof(TriRecord) * (size_t)n));
    CUDA_OK(cudaMalloc(&d_tris, sizeof(TriRecord) * (size_t)n));
    CUDA_OK(cudaMalloc(&d_tbox, sizeof(DBox) * (size_t)n));
    CUDA_OK(cudaMalloc(&d_bounds, sizeof(float) * 6));
    CUDA_OK(cudaMemcpy(d_bounds, binit, sizeof(binit), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&d_keys, 8 * (size_t)n));
    CUDA_OK(cudaMalloc(&d_keys2, 8 * (size_t)n));
    CUDA_OK(cudaMalloc(&d_idx, 4 * (size_t)n));
    CUDA_OK(cudaMalloc(&d_idx2, 4 * (size_t)n));
    k_lbvh_prepare<<<gb, 256>>>(d_verts, d_faces, n, d_recs, d_tbox, d_bounds);
    k_lbvh_morton<<<gb, 256>>>(d_tbox, n, d_bounds, d_keys, d_idx);
    CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, n, 0, 63));
    CUDA_OK(cudaMalloc(&d_tmp, tmp_bytes));
    CUDA_OK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, n, 0, 63));
    CUDA_OK(cudaMalloc(&N.child, sizeof(int2) * (size_t)n));
    CUDA_OK(cudaMalloc(&N.range, sizeof(int2) * (size_t)n));
    CUDA_OK(cudaMalloc(&N.parent, sizeof(int) * 2 * (size_t)n));
    CUDA_OK(cudaMalloc(&N.box, sizeof(DBox) * 2 * (size_t)n));
    CUDA_OK(cudaMalloc(&N.flag, sizeof(int) * (size_t)n));
    CUDA_OK(cudaMalloc(&N.cost, sizeof(float) * 8 * (size_t)n));
    CUDA_OK(cudaMalloc(&N.dec, sizeof(uint32_t) * (size_t)n));
    CUDA_OK(cudaMemset(N.flag, 0, sizeof(int) * (size_t)n));
    CUDA_OK(cudaMemset(N.parent, 0xFF, sizeof(int) * 2 * (size_t)n));
    if (n > 1) k_lbvh_hierarchy<<<gb, 256>>>(d_keys2, n, N);
#if IRIS_SAH_TREELET > 0
    if (n > 3 && g_sah_treelets) {
        // binned-SAH rebuild (bvh_device.cuh): top tree over the clusters (host, a few thousand boxes), then every cluster by one CTA
        int n_list = 0, C = 0;
        bool top_done = false;
        CUDA_OK(cudaMalloc(&d_list, sizeof(int) * (size_t)n));
        CUDA_OK(cudaMalloc(&d_nlist, sizeof(int)));
        CUDA_OK(cudaMemset(d_nlist, 0, sizeof(int)));
        if (g_sah_top && n > IRIS_SAH_TREELET) {
            CUDA_OK(cudaMalloc(&d_cl, sizeof(int2) * (size_t)n));
            k_lbvh_clusters<<<(unsigned)((2 * (size_t)n + 255) / 256), 256>>>(n, N, IRIS_SAH_TREELET, d_cl, d_nlist);
            CUDA_OK(cudaMemcpy(&C, d_nlist, sizeof(int), cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemset(d_nlist, 0, sizeof(int)));
        }
        if (C >= 2) {
            std::vector<int2> cl((size_t)C);
            std::vector<float> cbox((size_t)C * 6), sbox((size_t)C * 6);
            CUDA_OK(cudaMalloc(&d_cbox, sizeof(float) * 6 * (size_t)C));
            k_lbvh_cluster_boxes<<<(unsigned)(((size_t)C * 32 + 255) / 256), 256>>>(d_cl, C, d_tbox, d_idx2, d_cbox);
            CUDA_OK(cudaMemcpy(cl.data(), d_cl, sizeof(int2) * (size_t)C, cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemcpy(cbox.data(), d_cbox, sizeof(float) * 6 * (size_t)C, cudaMemcpyDeviceToHost));
            std::vector<int32_t> by_start((size_t)C), ssize((size_t)C);
            for (int c = 0; c < C; ++c) by_start[(size_t)c] = c;
            std::sort(by_start.begin(), by_start.end(), [&](int32_t x, int32_t y) { return cl[(size_t)x].x < cl[(size_t)y].x; });
            bool tiles = true;                                   // the clusters must tile [0, n) exactly
            int64_t pos = 0;
            for (int c = 0; c < C && tiles; ++c) {
                const int2 q = cl[(size_t)by_start[(size_t)c]];
                tiles = q.x == pos && q.y >= 1;
                pos += q.y;
                for (int k = 0; k < 6; ++k) sbox[6 * (size_t)c + k] = cbox[6 * (size_t)by_start[(size_t)c] + k];
                ssize[(size_t)c] = q.y;
            }
            tiles = tiles && pos == n;
            if (tiles) {
                std::vector<int32_t> order((size_t)C), tf((size_t)C), tc((size_t)C), tl((size_t)C), tL((size_t)C), tR((size_t)C);
                const int n_top = host_sah_top(sbox.data(), ssize.data(), C, order.data(), tf.data(), tc.data(), tl.data(), tL.data(), tR.data());
                std::vector<int64_t> P((size_t)C + 1, 0);        // first sorted position of the cluster at tree-order position k
                std::vector<int3> perm((size_t)C), roots;
                for (int k = 0; k < C; ++k) {
                    const int2 q = cl[(size_t)by_start[(size_t)order[(size_t)k]]];
                    perm[(size_t)k] = make_int3(q.x, q.y, (int)P[(size_t)k]);
                    P[(size_t)k + 1] = P[(size_t)k] + q.y;
                }
                std::vector<int> rec;
                rec.reserve((size_t)5 * (size_t)n_top);
                std::vector<std::pair<int32_t, int32_t>> st;      // (top node, its id in the binary tree)
                st.push_back({0, 0});
                while (!st.empty()) {
                    const int32_t k = st.back().first, id = st.back().second;
                    st.pop_back();
                    const int32_t f0 = tf[(size_t)k], mid = f0 + tl[(size_t)k], end = f0 + tc[(size_t)k];
                    const int64_t a0 = P[(size_t)f0], am = P[(size_t)mid], ae = P[(size_t)end];
                    const int left_id = (am - a0 == 1) ? (int)(n - 1 + a0) : (int)(am - 1);
                    const int right_id = (ae - am == 1) ? (int)(n - 1 + am) : (int)am;
                    rec.insert(rec.end(), {(int)id, left_id, right_id, (int)a0, (int)(ae - 1)});
                    if (tL[(size_t)k] >= 0) st.push_back({tL[(size_t)k], left_id});
                    else if (am - a0 >= 2) roots.push_back(make_int3(left_id, (int)a0, (int)(am - 1)));
                    if (tR[(size_t)k] >= 0) st.push_back({tR[(size_t)k], right_id});
                    else if (ae - am >= 2) roots.push_back(make_int3(right_id, (int)am, (int)(ae - 1)));
                }
                std::vector<int> root_ids(roots.size());
                for (size_t k = 0; k < roots.size(); ++k) root_ids[k] = roots[k].x;
                CUDA_OK(cudaMalloc(&d_perm, sizeof(int3) * (size_t)C));
                CUDA_OK(cudaMalloc(&d_rec, sizeof(int) * std::max<size_t>(rec.size(), 1)));
                CUDA_OK(cudaMalloc(&d_roots, sizeof(int3) * std::max<size_t>(roots.size(), 1)));
                CUDA_OK(cudaMemcpy(d_perm, perm.data(), sizeof(int3) * (size_t)C, cudaMemcpyHostToDevice));
                CUDA_OK(cudaMemcpy(d_rec, rec.data(), sizeof(int) * rec.size(), cudaMemcpyHostToDevice));
                CUDA_OK(cudaMemcpy(d_roots, roots.data(), sizeof(int3) * roots.size(), cudaMemcpyHostToDevice));
                CUDA_OK(cudaMemcpy(d_list, root_ids.data(), sizeof(int) * root_ids.size(), cudaMemcpyHostToDevice));
                k_lbvh_permute<<<(unsigned)C, 128>>>(d_perm, d_idx2, d_idx);
                std::swap(d_idx, d_idx2);                         // d_idx2 stays "the sorted order" for everything below
                const int n_w = (int)std::max(rec.size() / 5, roots.size());
                k_lbvh_write_top<<<(unsigned)((n_w + 255) / 256), 256>>>(d_rec, (int)(rec.size() / 5), d_roots, (int)roots.size(), N);
                n_list = (int)root_ids.size();
                top_done = true;
            }
        }
        if (!top_done) {
            k_lbvh_treelet_roots<<<gb, 256>>>(n, N, IRIS_SAH_TREELET, d_list, d_nlist);
            CUDA_OK(cudaMemcpy(&n_list, d_nlist, sizeof(int), cudaMemcpyDeviceToHost));
        }
        if (n_list > 0) {
            CUDA_OK(cudaFuncSetAttribute(k_lbvh_sah_treelets<IRIS_SAH_TREELET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SAH_SMEM_BYTES(IRIS_SAH_TREELET)));
            k_lbvh_sah_treelets<IRIS_SAH_TREELET><<<(unsigned)std::min(n_list, 148 * 16), SAH_THREADS, SAH_SMEM_BYTES(IRIS_SAH_TREELET)>>>(d_list, n_list, n, N, d_tbox, d_idx2);
        }
    }
#endif
    k_lbvh_fit<<<gb, 256>>>(d_tbox, d_idx2, n, d_bounds, N);
    CUDA_OK(cudaMalloc(&d_wide, sizeof(Bvh8Node) * (size_t)n));
    CUDA_OK(cudaMalloc(&d_wide_bin, sizeof(int) * (size_t)n));
    CUDA_OK(cudaMalloc(&d_wide_depth, sizeof(int) * (size_t)n));
    CUDA_OK(cudaMalloc(&d_counters, sizeof(int) * 3));
    CUDA_OK(cudaMemcpy(d_counters, counters, sizeof(counters), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemset(d_wide_bin, 0, sizeof(int)));           // wide node 0 <-> binary root (node 0)
    {
        const int one = 1;
        CUDA_OK(cudaMemcpy(d_wide_depth, &one, sizeof(int), cudaMemcpyHostToDevice));
    }
    for (int begin = 0, end = 1; begin < end;) {                 // one launch per level of the wide tree
        k_lbvh_collapse<<<(unsigned)((end - begin + 127) / 128), 128>>>(begin, end, n, N, d_recs, d_idx2, d_wide_bin, d_wide_depth, d_wide, d_tris, d_counters);
        CUDA_OK(cudaMemcpy(counters, d_counters, sizeof(counters), cudaMemcpyDeviceToHost));
        begin = end;
        end = counters[0];
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMalloc(&s->nodes, sizeof(Bvh8Node) * (size_t)counters[0]));
    CUDA_OK(cudaMemcpy(s->nodes, d_wide, sizeof(Bvh8Node) * (size_t)counters[0], cudaMemcpyDeviceToDevice));
    s->tris = reinterpret_cast<float4 *>(d_tris);
    d_tris = nullptr;
    s->stats.n_tris = n;
    s->stats.n_nodes = counters[0];
    s->stats.node_bytes = (int64_t)sizeof(Bvh8Node) * counters[0];
    s->stats.tri_bytes = (int64_t)sizeof(TriRecord) * n;
    s->stats.max_depth = counters[2];
    s->stats.sah_cost = 0.f;
    for (int k = 0; k < 3; ++k) { s->stats.bounds_lo[k] = hb[k]; s->stats.bounds_hi[k] = hb[3 + k]; }
    if (counters[1] != n) e = cudaErrorUnknown;                  // every triangle must have been emitted exactly once
done:
    cudaFree(d_verts); cudaFree(d_faces); cudaFree(d_recs); cudaFree(d_tris); cudaFree(d_tbox); cudaFree(d_bounds);
    cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_idx); cudaFree(d_idx2); cudaFree(d_tmp);
    cudaFree(N.child); cudaFree(N.range); cudaFree(N.parent); cudaFree(N.box); cudaFree(N.flag); cudaFree(N.cost); cudaFree(N.dec);
    cudaFree(d_wide); cudaFree(d_wide_bin); cudaFree(d_wide_depth); cudaFree(d_counters); cudaFree(d_list); cudaFree(d_nlist); cudaFree(d_rec); cudaFree(d_cl); cudaFree(d_perm); cudaFree(d_roots); cudaFree(d_cbox);
    return e;
}

extern "C" {

const char *iris_last_error(void) { return g_err.c_str(); }
const char *iris_version(void) { return "iris_b200 0.2 sm_100a"; }
int64_t iris_abi_info(int what) {
    switch (what) {
    case 0: return IRIS_ABI_VERSION;
    case 1: return (int64_t)sizeof(IrisShadeParams);
    case 2: return (int64_t)sizeof(IrisSampler);
    case 3: return (int64_t)sizeof(IrisSceneStats);
    default: return -1;
    }
}
int64_t iris_launch_count(void) { return g_launches.load(); }

int iris_set_option(const char *name, int value) {
    if (name && std::strcmp(name, "field_forward_impl") == 0 && (value == 0 || value == 1)) { g_field_impl = value; return IRIS_OK; }
    if (name && std::strcmp(name, "intersect_impl") == 0 && (value == 0 || value == 1)) { g_intersect_impl = value; return IRIS_OK; }
    if (name && std::strcmp(name, "trace_smem_carveout_pct") == 0 && value >= -1 && value <= 100) {
        // shared-memory share of the SM's 228 KB for the tracing kernels (they use none, or 14 KB for the bake sort): -1 = driver default, 0 = all L1
        CUDA_TRY(cudaFuncSetAttribute(k_trace_queue, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_intersect, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_intersect_persistent, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_primary, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_bake<0>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_bake<1>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        return IRIS_OK;
    }
    if (name && std::strcmp(name, "field_backward_impl") == 0 && value >= 0 && value <= 2) { g_field_bwd_impl = value; return IRIS_OK; }
    if (name && std::strcmp(name, "wave_impl") == 0 && (value == 0 || value == 1)) { g_wave_impl = value; return IRIS_OK; }
    if (name && std::strcmp(name, "wave_compact") == 0 && (value == 0 || value == 1)) { g_wave_compact = value; return IRIS_OK; }
    if (name && std::strcmp(name, "bake_impl") == 0 && value >= 0 && value <= 2) { g_bake_impl = value; return IRIS_OK; }
    if (name && std::strcmp(name, "single_impl") == 0 && (value == 0 || value == 1)) { g_single_impl = value; return IRIS_OK; }
    if (name && std::strcmp(name, "single_chunk_log2") == 0 && value >= 10 && value <= 30) { g_single_chunk = (int64_t)1 << value; return IRIS_OK; }
    if (name && std::strcmp(name, "persist_ctas_per_sm") == 0 && value >= 1 && value <= 16) { g_persist_ctas = value; return IRIS_OK; }
    if (name && std::strcmp(name, "lbvh_sah_top") == 0 && (value == 0 || value == 1)) { g_sah_top = value; return IRIS_OK; }
    if (name && std::strcmp(name, "lbvh_sah_treelets") == 0 && (value == 0 || value == 1)) { g_sah_treelets = value; return IRIS_OK; }
    if (name && std::strcmp(name, "tc5_bwd_ctas_per_sm") == 0 && value >= 1 && value <= 2) { g_tc5_bwd_ctas = value; return IRIS_OK; }
    if (name && std::strcmp(name, "scatter_ctas_per_sm") == 0 && value >= 0 && value <= 8) { g_scatter_ctas = value; return IRIS_OK; }
    if (name && std::strcmp(name, "tc5_ctas_per_sm") == 0 && value >= 1 && value <= 8) { g_tc5_ctas = value; return IRIS_OK; }
    if (name && std::strcmp(name, "field_smem_carveout_pct") == 0 && value >= 0 && value <= 100) {
        // how much of the SM's 228 KB the field kernels ask to be shared memory: the rest is L1, which the hash-grid gathers live on
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward_tc5<true>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward_tc5<false>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward<true>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward<false>, cudaFuncAttributePreferredSharedMemoryCarveout, value));
        return IRIS_OK;
    }
    if (name && std::strcmp(name, "tc5_debug") == 0) { CUDA_TRY(cudaMemcpyToSymbol(g_tc5_debug, &value, sizeof(int))); return IRIS_OK; }
    return fail(IRIS_ERR_INVALID, "unknown option or value");
}

int iris_profile_enable(int on) {
    g_prof_on = on != 0;
    return IRIS_OK;
}
const char *iris_profile_name(int kernel_id) { return kernel_id >= 0 && kernel_id < K_COUNT ? g_kernel_names[kernel_id] : nullptr; }
int iris_profile_read(int kernel_id, int64_t *launches, double *total_ms, int reset) {
    if (kernel_id < 0 || kernel_id >= K_COUNT) return fail(IRIS_ERR_INVALID, "bad kernel id");
    for (auto &sp : g_spans) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(sp.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, sp.a, sp.b);
        if (e != cudaSuccess) return fail(IRIS_ERR_CUDA, std::string("profile: ") + cudaGetErrorString(e));
        g_prof_ms[sp.id] += ms;
        g_prof_n[sp.id] += 1;
        g_free_events.push_back(sp.a);
        g_free_events.push_back(sp.b);
    }
    g_spans.clear();
    if (launches) *launches = g_prof_n[kernel_id];
    if (total_ms) *total_ms = g_prof_ms[kernel_id];
    if (reset) { g_prof_n[kernel_id] = 0; g_prof_ms[kernel_id] = 0.0; }
    return IRIS_OK;
}

int64_t iris_field_levels(float *scale, uint32_t *res, uint32_t *size, uint32_t *offset) {
    FieldLevel lv[FIELD_LEVELS];
    const int64_t total = field_level_table(lv);
    for (int l = 0; l < FIELD_LEVELS; ++l) {
        if (scale) scale[l] = lv[l].scale;
        if (res) res[l] = lv[l].res;
        if (size) size[l] = lv[l].size;
        if (offset) offset[l] = lv[l].offset;
    }
    return total;
}

int iris_scene_create(const float *verts, int64_t n_verts, const int32_t *faces, int64_t n_faces, int device, int builder, IrisScene **out) {
    if (!out) return fail(IRIS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_faces < 0 || n_verts < 0 || (n_faces > 0 && (!verts || !faces))) return fail(IRIS_ERR_INVALID, "bad mesh arguments");
    if (n_faces > 0x7FFFFFF0ll / 3) return fail(IRIS_ERR_INVALID, "too many faces");
    if (builder != 0 && builder != 1) return fail(IRIS_ERR_INVALID, "builder must be 0 (host binned SAH) or 1 (on-device builder)");
    for (int64_t i = 0; i < 3 * n_faces; ++i)
        if (faces[i] < 0 || faces[i] >= n_verts) return fail(IRIS_ERR_INVALID, "face index out of range");
    CUDA_TRY(cudaSetDevice(device));
    int rc = ensure_device_setup(device);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    if (builder == 1 && n_faces > 0) {
        IrisScene *s = new IrisScene();
        s->device = device;
        cudaError_t e = device_bvh_build(verts, n_verts, faces, n_faces, s);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess || 2 * s->stats.max_depth > IRIS_STACK) {
            const std::string msg = e != cudaSuccess ? std::string("device BVH build: ") + cudaGetErrorString(e)
                                                     : "device BVH deeper than the traversal stack (" + std::to_string(s->stats.max_depth) + " levels)";
            cudaFree(s->nodes);
            cudaFree(s->tris);
            delete s;
            return fail(e != cudaSuccess ? IRIS_ERR_CUDA : IRIS_ERR_INVALID, msg);
        }
        s->stats.build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *out = s;
        return IRIS_OK;
    }
    HostBvh hb;
    if (host_bvh_build(verts, n_verts, faces, n_faces, &hb) != 0) return fail(IRIS_ERR_NOMEM, "host BVH build failed");
    if (2 * hb.max_depth > IRIS_STACK) {
        host_bvh_free(&hb);
        return fail(IRIS_ERR_INVALID, "BVH deeper than the traversal stack (" + std::to_string(hb.max_depth) + " levels)");
    }
    IrisScene *s = new IrisScene();
    s->device = device;
    const size_t nb = sizeof(Bvh8Node) * (size_t)hb.n_nodes, tb = sizeof(TriRecord) * (size_t)(hb.n_tris > 0 ? hb.n_tris : 1);
    cudaError_t e = cudaMalloc(&s->nodes, nb);
    if (e == cudaSuccess) e = cudaMalloc(&s->tris, tb);
    if (e == cudaSuccess) e = cudaMemcpy(s->nodes, hb.nodes, nb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && hb.n_tris > 0) e = cudaMemcpy(s->tris, hb.tris, sizeof(TriRecord) * (size_t)hb.n_tris, cudaMemcpyHostToDevice);
    s->stats.n_tris = hb.n_tris;
    s->stats.n_nodes = hb.n_nodes;
    s->stats.node_bytes = (int64_t)nb;
    s->stats.tri_bytes = (int64_t)(sizeof(TriRecord) * (size_t)hb.n_tris);
    s->stats.sah_cost = hb.sah_cost;
    s->stats.max_depth = hb.max_depth;
    for (int k = 0; k < 3; ++k) { s->stats.bounds_lo[k] = hb.lo[k]; s->stats.bounds_hi[k] = hb.hi[k]; }
    host_bvh_free(&hb);
    if (e != cudaSuccess) {
        cudaFree(s->nodes);
        cudaFree(s->tris);
        delete s;
        return fail(IRIS_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e));
    }
    s->stats.build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = s;
    return IRIS_OK;
}

void iris_scene_destroy(IrisScene *s) {
    if (!s) return;
    cudaFree(s->nodes);
    cudaFree(s->tris);
    delete s;
}

int iris_scene_stats(const IrisScene *s, IrisSceneStats *out) {
    if (!s || !out) return fail(IRIS_ERR_INVALID, "NULL argument");
    *out = s->stats;
    return IRIS_OK;
}

// next slot of the counter ring, zeroed on the caller's stream
static std::atomic<unsigned> g_counter_slot{0};
static int next_ray_counter(unsigned long long **out, cudaStream_t st) {
    void *base = nullptr;
    CUDA_TRY(cudaGetSymbolAddress(&base, g_ray_counters));
    *out = reinterpret_cast<unsigned long long *>(base) + (g_counter_slot.fetch_add(1) % IRIS_COUNTER_RING);
    CUDA_TRY(cudaMemsetAsync(*out, 0, sizeof(unsigned long long), st));
    return IRIS_OK;
}

int iris_intersect(const IrisScene *s, const float *o, const float *d, int64_t n, float *t, int32_t *prim, float *uv, float *p, float *nrm,
                   void *stream) {
    if (!s) return fail(IRIS_ERR_INVALID, "scene is NULL");
    if (n < 0) return fail(IRIS_ERR_INVALID, "n < 0");
    if (n == 0) return IRIS_OK;
    if (!o || !d) return fail(IRIS_ERR_INVALID, "ray arrays are NULL");
    {
        ProfScope ps(K_INTERSECT, (cudaStream_t)stream);
        if (g_intersect_impl == 1) {
            unsigned long long *cnt = nullptr;
            int rc = next_ray_counter(&cnt, (cudaStream_t)stream);
            if (rc) return rc;
            const int64_t want = (n + IRIS_BLOCK - 1) / IRIS_BLOCK;
            const int grid = (int)std::min<int64_t>(want, (int64_t)g_persist_ctas * (g_sm_count > 0 ? g_sm_count : 148));
            k_intersect_persistent<<<grid, IRIS_BLOCK, 0, (cudaStream_t)stream>>>(view_of(s), o, d, n, t, prim, uv, p, nrm, cnt);
        } else {
            k_intersect<<<blocks_for(n), IRIS_BLOCK, 0, (cudaStream_t)stream>>>(view_of(s), o, d, n, t, prim, uv, p, nrm);
        }
    }
    LAUNCHED();
    return IRIS_OK;
}

__global__ void k_sampler_fill(IrisSampler smp, int64_t n, int dims, float *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int b = 0; 4 * b < dims; ++b) {
        const float4 u = sample4(smp, i, b);
        const float v[4] = {u.x, u.y, u.z, u.w};
        for (int k = 0; k < 4 && 4 * b + k < dims; ++k) out[i * dims + 4 * b + k] = v[k];
    }
}

int iris_sampler_fill(uint64_t seed, uint64_t lane_offset, int64_t n, int32_t dims, float *out, void *stream) {
    if (n < 0 || dims <= 0 || !out) return fail(IRIS_ERR_INVALID, "bad sampler_fill arguments");
    if (n == 0) return IRIS_OK;
    IrisSampler smp{nullptr, 0, seed, lane_offset};
    k_sampler_fill<<<blocks_for(n), IRIS_BLOCK, 0, (cudaStream_t)stream>>>(smp, n, dims, out);
    LAUNCHED();
    return IRIS_OK;
}

static int check_params(const IrisShadeParams *P, bool need_field) {
    if (!P) return fail(IRIS_ERR_INVALID, "params is NULL");
    if (!P->emitter_of_face || !P->radiance || P->n_emitters <= 0 || !P->face_of_emitter || !P->emitter_vertices || !P->emitter_area ||
        !P->emitter_pdf || !P->emitter_cdf)
        return fail(IRIS_ERR_INVALID, "emitter tables missing");
    if (!P->slf_inds || !P->slf_radiance || P->slf_H <= 0 || !(P->slf_range > 0.f)) return fail(IRIS_ERR_INVALID, "SLF tables missing");
    if (need_field && (!P->grid_f16 || !P->mlp_f16 || !(P->field_range > 0.f))) return fail(IRIS_ERR_INVALID, "BRDF field tables missing");
    return IRIS_OK;
}

// ray queue of one chunk of the wavefront bounce: counters | rays (2 x 2 float4) | hits (2 float4) | state streams
static const int64_t SINGLE_MAX_CHUNKS = 1024;
static int64_t single_chunk_samples(int64_t n) {
    int64_t c = std::min<int64_t>(n, g_single_chunk);
    while ((n + c - 1) / c > SINGLE_MAX_CHUNKS) c *= 2;
    return std::max<int64_t>(c, 1);
}
static int64_t single_queue_bytes(int64_t n) {
    return 8 * SINGLE_MAX_CHUNKS + (4 + 2 + IRIS_SINGLE_STATE_STREAMS) * 16 * single_chunk_samples(n);
}

int iris_bake(const IrisScene *s, const IrisShadeParams *P, int mode, float roughness, const float *position, const float *normal, const float *wo,
              int64_t n_pixels, int32_t spp, const IrisSampler *sampler, float *out0, float *out1, void *stream) {
    if (!s) return fail(IRIS_ERR_INVALID, "scene is NULL");
    int rc = check_params(P, false);
    if (rc) return rc;
    if (n_pixels < 0 || spp <= 0 || !sampler) return fail(IRIS_ERR_INVALID, "bad bake arguments");
    if (mode != 0 && mode != 1) return fail(IRIS_ERR_INVALID, "mode must be 0 (diffuse) or 1 (specular)");
    if (n_pixels == 0) return IRIS_OK;
    if (!position || !normal || !out0 || (mode == 1 && (!wo || !out1))) return fail(IRIS_ERR_INVALID, "NULL array");
    if (sampler->U && sampler->stride < 2) return fail(IRIS_ERR_INVALID, "sampler stride < 2");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(out0, 0, sizeof(float) * 3 * (size_t)n_pixels, st));
    if (mode == 1) CUDA_TRY(cudaMemsetAsync(out1, 0, sizeof(float) * 3 * (size_t)n_pixels, st));
    const int64_t n = n_pixels * spp;
    if (g_bake_impl == 2) {           // persistent warps, generator and radiance lookup in the kernel
        unsigned long long *cnt = nullptr;
        if ((rc = next_ray_counter(&cnt, st))) return rc;
        ProfScope ps(mode == 0 ? K_BAKE_DIFFUSE : K_BAKE_SPECULAR, st);
        const int grid = (int)std::min<int64_t>(blocks_for(n), (int64_t)g_persist_ctas * (g_sm_count > 0 ? g_sm_count : 148));
        if (mode == 0) k_bake_persistent<0><<<grid, IRIS_BLOCK, 0, st>>>(view_of(s), *P, *sampler, roughness, position, normal, wo, n_pixels, spp, out0, out1, cnt);
        else k_bake_persistent<1><<<grid, IRIS_BLOCK, 0, st>>>(view_of(s), *P, *sampler, roughness, position, normal, wo, n_pixels, spp, out0, out1, cnt);
        LAUNCHED();
        return IRIS_OK;
    }
    if (g_bake_impl == 0) {
        ProfScope ps(mode == 0 ? K_BAKE_DIFFUSE : K_BAKE_SPECULAR, st);
        const unsigned gb = (unsigned)((n + IRIS_SORT_BLOCK - 1) / IRIS_SORT_BLOCK);
        if (mode == 0) k_bake<0><<<gb, IRIS_SORT_BLOCK, 0, st>>>(view_of(s), *P, *sampler, roughness, position, normal, wo, n_pixels, spp, out0, out1);
        else k_bake<1><<<gb, IRIS_SORT_BLOCK, 0, st>>>(view_of(s), *P, *sampler, roughness, position, normal, wo, n_pixels, spp, out0, out1);
        LAUNCHED();
        return IRIS_OK;
    }
    // ray-queue form: generate -> persistent trace -> shade, in chunks; the queue is stream-ordered scratch memory
    const int64_t nc_max = single_chunk_samples(n);
    const int64_t n_chunks = (n + nc_max - 1) / nc_max;
    unsigned char *scratch = nullptr;
    CUDA_TRY(cudaMallocAsync(&scratch, 8 * SINGLE_MAX_CHUNKS + 3 * 16 * (size_t)nc_max, st));
    unsigned long long *counters = reinterpret_cast<unsigned long long *>(scratch);
    float4 *ro = reinterpret_cast<float4 *>(scratch + 8 * SINGLE_MAX_CHUNKS), *rd = ro + nc_max, *hit = rd + nc_max;
    CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * (size_t)n_chunks, st));
    int chunk = 0;
    for (int64_t i0 = 0; i0 < n; i0 += nc_max, ++chunk) {
        const int64_t nc = std::min(nc_max, n - i0);
        {
            ProfScope ps(K_BAKE_GEN, st);
            if (mode == 0) k_bake_gen<0><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(*sampler, roughness, position, normal, wo, i0, nc, spp, ro, rd);
            else k_bake_gen<1><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(*sampler, roughness, position, normal, wo, i0, nc, spp, ro, rd);
        }
        LAUNCHED();
        {
            ProfScope ps(K_TRACE_QUEUE, st);
            const int grid = (int)std::min<int64_t>(blocks_for(nc), (int64_t)g_persist_ctas * (g_sm_count > 0 ? g_sm_count : 148));
            k_trace_queue<<<grid, IRIS_BLOCK, 0, st>>>(view_of(s), ro, rd, nc, 0, hit, counters + chunk, nullptr, 0);
        }
        LAUNCHED();
        {
            ProfScope ps(K_BAKE_SHADE, st);
            if (mode == 0) k_bake_shade<0><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(view_of(s), *P, roughness, normal, wo, i0, nc, spp, rd, hit, out0, out1);
            else k_bake_shade<1><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(view_of(s), *P, roughness, normal, wo, i0, nc, spp, rd, hit, out0, out1);
        }
        LAUNCHED();
    }
    CUDA_TRY(cudaFreeAsync(scratch, st));
    return IRIS_OK;
}

static int launch_field(const IrisShadeParams *P, int64_t n, const float *position, float *mat, const float4 *w0, float4 *w1, float4 *w2, cudaStream_t st,
                        __half *x_save = nullptr, int pair = 1, const unsigned long long *n_dev = nullptr) {
    static bool attr_done_dev[64] = {false};       // function attributes belong to the device context: set them once per device
    int cur_dev = 0;
    CUDA_TRY(cudaGetDevice(&cur_dev));
    bool &attr_done = attr_done_dev[cur_dev & 63];
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIELD_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIELD_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward_tc5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC5_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward_tc5<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC5_SMEM_BYTES));
        // The encoder lives on L1: with the driver's default carveout for the tcgen05 kernel (all shared memory, no L1) it runs 2.4x
        // slower.  70% of the SM's 228 KB as shared memory holds 4 CTAs and leaves ~68 KB of L1 for the hash-grid gathers
        // (measured: 2.96 G samples/s on camera-coherent positions, 1.33 G/s on uniformly random ones; 35% favours the latter).
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 70));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 70));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward_tc5<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 70));
        CUDA_TRY(cudaFuncSetAttribute(k_field_forward_tc5<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 70));
        attr_done = true;
    }
    const int64_t tiles = (n + IRIS_BLOCK - 1) / IRIS_BLOCK;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)g_sm_count * 4);
    ProfScope ps(K_FIELD_FORWARD, st);
    if (g_field_impl == 1) {
        const unsigned g5 = (unsigned)std::min<int64_t>(tiles, (int64_t)g_sm_count * g_tc5_ctas);
        if (w0) k_field_forward_tc5<true><<<g5, TC5_ROWS, TC5_SMEM_BYTES, st>>>(*P, n, position, mat, w0, w1, w2, x_save, pair, n_dev);
        else k_field_forward_tc5<false><<<g5, TC5_ROWS, TC5_SMEM_BYTES, st>>>(*P, n, position, mat, w0, w1, w2, x_save, pair, n_dev);
    } else {
        if (w0) k_field_forward<true><<<grid, IRIS_BLOCK, FIELD_SMEM_BYTES, st>>>(*P, n, position, mat, w0, w1, w2, x_save, pair, n_dev);
        else k_field_forward<false><<<grid, IRIS_BLOCK, FIELD_SMEM_BYTES, st>>>(*P, n, position, mat, w0, w1, w2, x_save, pair, n_dev);
    }
    LAUNCHED();
    return IRIS_OK;
}

int iris_field_forward(const IrisShadeParams *P, const float *position, int64_t n, float *mat, void *encoded, void *stream) {
    if (!P || !P->grid_f16 || !P->mlp_f16 || !(P->field_range > 0.f)) return fail(IRIS_ERR_INVALID, "BRDF field tables missing");
    if (n < 0) return fail(IRIS_ERR_INVALID, "n < 0");
    if (n == 0) return IRIS_OK;
    if (!position || !mat) return fail(IRIS_ERR_INVALID, "NULL array");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ensure_device_setup(dev);
    if (rc) return rc;
    if (reinterpret_cast<uintptr_t>(encoded) & 15) return fail(IRIS_ERR_INVALID, "encoded must be 16-byte aligned");
    return launch_field(P, n, position, mat, nullptr, nullptr, nullptr, (cudaStream_t)stream, reinterpret_cast<__half *>(encoded));
}

#ifndef FIELD_BWD_CHUNK_LOG2
#define FIELD_BWD_CHUNK_LOG2 23
#endif
#define FIELD_BWD_CHUNK (1ll << FIELD_BWD_CHUNK_LOG2)
static int run_field_backward(const IrisShadeParams *P, int64_t n, const float *position, const float4 *r5, const float *d_mat, float *d_params,
                              void *workspace, int64_t workspace_bytes, cudaStream_t st, __half *x_saved = nullptr) {
    static bool attr_done_dev[64] = {false};       // function attributes belong to the device context: set them once per device
    int cur_dev = 0;
    CUDA_TRY(cudaGetDevice(&cur_dev));
    bool &attr_done = attr_done_dev[cur_dev & 63];
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(k_field_backward_dgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIELD_BWD_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_field_backward_dgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIELD_BWD_SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(k_field_backward_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, FIELD_WGRAD_SMEM_BYTES));
        attr_done = true;
    }
    const int64_t chunk = std::min<int64_t>(n, FIELD_BWD_CHUNK);
    if (workspace_bytes < chunk * FIELD_ACT_BYTES_PER_SAMPLE) return fail(IRIS_ERR_WORKSPACE, "field backward: workspace too small");
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        const int64_t m = std::min<int64_t>(chunk, n - c0);
        FieldAct act = field_act_carve(workspace, m);
        if (x_saved) act.X = x_saved + 64 * c0;        // encoded inputs kept by the forward pass: not recomputed, not copied
        const int64_t tiles = (m + FIELD_BWD_BLOCK - 1) / FIELD_BWD_BLOCK;
        const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)g_sm_count * (512 / FIELD_BWD_BLOCK));
        const bool fused = g_field_bwd_impl >= 1 && x_saved != nullptr;          // one-kernel dgrad + wgrad on tcgen05 (field_bwd_tc5.cuh / field_bwd_tc5v2.cuh)
        if (fused && g_field_bwd_impl == 2) {
            static bool fattr2[64] = {false};
            if (!fattr2[cur_dev & 63]) {
                CUDA_TRY(cudaFuncSetAttribute(k_field_backward_tc5v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT6_SMEM_BYTES));
                CUDA_TRY(cudaFuncSetAttribute(k_field_backward_tc5v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT6_SMEM_BYTES));
                fattr2[cur_dev & 63] = true;
            }
            CUtensorMap tm_x, tm_dx;
            int rc2 = make_row_tile_map(&tm_x, act.X, m);
            if (rc2 == IRIS_OK) rc2 = make_row_tile_map(&tm_dx, act.dx, m);
            if (rc2) return rc2;
            ProfScope ps(K_FIELD_BACKWARD_TC5, st);
            const unsigned gf = (unsigned)std::min<int64_t>((m + TC5_ROWS - 1) / TC5_ROWS, (int64_t)g_sm_count * g_tc5_bwd_ctas);
            if (r5) k_field_backward_tc5v2<true><<<gf, BT6_THREADS, BT6_SMEM_BYTES, st>>>(tm_x, tm_dx, *P, m, r5 + c0, d_mat + 5 * c0, act.s, d_params);
            else k_field_backward_tc5v2<false><<<gf, BT6_THREADS, BT6_SMEM_BYTES, st>>>(tm_x, tm_dx, *P, m, nullptr, d_mat + 5 * c0, act.s, d_params);
        } else if (fused) {
            static bool fattr[64] = {false};
            if (!fattr[cur_dev & 63]) {
                CUDA_TRY(cudaFuncSetAttribute(k_field_backward_tc5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT5_SMEM_BYTES));
                CUDA_TRY(cudaFuncSetAttribute(k_field_backward_tc5<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BT5_SMEM_BYTES));
                fattr[cur_dev & 63] = true;
            }
            ProfScope ps(K_FIELD_BACKWARD_TC5, st);
            const unsigned gf = (unsigned)std::min<int64_t>((m + TC5_ROWS - 1) / TC5_ROWS, (int64_t)g_sm_count * g_tc5_bwd_ctas);
            if (r5) k_field_backward_tc5<true><<<gf, TC5_ROWS, BT5_SMEM_BYTES, st>>>(*P, m, r5 + c0, d_mat + 5 * c0, act.X, act.dx, act.s, d_params);
            else k_field_backward_tc5<false><<<gf, TC5_ROWS, BT5_SMEM_BYTES, st>>>(*P, m, nullptr, d_mat + 5 * c0, act.X, act.dx, act.s, d_params);
        } else {
            ProfScope ps(K_FIELD_BACKWARD, st);
            if (r5) k_field_backward_dgrad<true><<<grid, FIELD_BWD_BLOCK, FIELD_BWD_SMEM_BYTES, st>>>(*P, m, nullptr, r5 + c0, d_mat + 5 * c0, act, x_saved ? 1 : 0);
            else k_field_backward_dgrad<false><<<grid, FIELD_BWD_BLOCK, FIELD_BWD_SMEM_BYTES, st>>>(*P, m, position + 3 * c0, nullptr, d_mat + 5 * c0, act, x_saved ? 1 : 0);
        }
        LAUNCHED();
        {
            ProfScope ps(K_FIELD_SCATTER, st);
            unsigned gs = (unsigned)((m + 255) / 256);
            if (g_scatter_ctas > 0) gs = std::min<unsigned>(gs, (unsigned)(g_scatter_ctas * (g_sm_count > 0 ? g_sm_count : 148)));
            if (r5) k_field_backward_scatter<true><<<gs, 256, 0, st>>>(*P, m, nullptr, r5 + c0, act, d_params + 9216);
            else k_field_backward_scatter<false><<<gs, 256, 0, st>>>(*P, m, position + 3 * c0, nullptr, act, d_params + 9216);
        }
        LAUNCHED();
        if (!fused) {
            ProfScope ps(K_FIELD_WGRAD, st);
            const unsigned g2 = (unsigned)std::min<int64_t>((m + 31) / 32, (int64_t)g_sm_count * 3);
            k_field_backward_wgrad<<<g2, IRIS_BLOCK, FIELD_WGRAD_SMEM_BYTES, st>>>(act, m, d_params);
            LAUNCHED();
        }
    }
    return IRIS_OK;
}

int64_t iris_field_backward_workspace_bytes(int64_t n) { return std::min<int64_t>(std::max<int64_t>(n, 1), FIELD_BWD_CHUNK) * FIELD_ACT_BYTES_PER_SAMPLE; }

int iris_field_backward(const IrisShadeParams *P, const float *position, const float *d_mat, int64_t n, float *d_params, const void *encoded,
                        void *workspace, int64_t workspace_bytes, void *stream) {
    if (!P || !P->grid_f16 || !P->mlp_f16 || !(P->field_range > 0.f)) return fail(IRIS_ERR_INVALID, "BRDF field tables missing");
    if (n < 0) return fail(IRIS_ERR_INVALID, "n < 0");
    if (n == 0) return IRIS_OK;
    if (!position || !d_mat || !d_params || !workspace) return fail(IRIS_ERR_INVALID, "NULL array");
    if (reinterpret_cast<uintptr_t>(workspace) & 15) return fail(IRIS_ERR_INVALID, "workspace must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_params) & 15) return fail(IRIS_ERR_INVALID, "d_params must be 16-byte aligned (16-byte vector reductions into the grid gradient)");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ensure_device_setup(dev);
    if (rc) return rc;
    if (reinterpret_cast<uintptr_t>(encoded) & 15) return fail(IRIS_ERR_INVALID, "encoded must be 16-byte aligned");
    return run_field_backward(P, n, position, nullptr, d_mat, d_params, workspace, workspace_bytes, (cudaStream_t)stream,
                              reinterpret_cast<__half *>(const_cast<void *>(encoded)));
}

int64_t iris_single_workspace_bytes(int64_t n_pixels, int32_t spp) {
    const int64_t n = n_pixels * (int64_t)spp;
    const int64_t fwd = 3 * 16 * n + single_queue_bytes(n);                          // w0,w1,w2 | ray queue of one chunk
    const int64_t bwd = ((5 * 4 * n + 15) / 16) * 16 + iris_field_backward_workspace_bytes(n);   // d_mat | activation streams of one chunk
    return std::max(fwd, bwd);
}
// per sample: 6 float4 of estimator record; separately (optional) the 64 fp16 encoded field inputs the field adjoint reads back
int64_t iris_single_record_bytes(int64_t n_pixels, int32_t spp) { return 6 * 16 * n_pixels * (int64_t)spp; }
int64_t iris_single_encoded_bytes(int64_t n_pixels, int32_t spp) { return 128 * n_pixels * (int64_t)spp; }

int iris_single_forward(const IrisScene *s, const IrisShadeParams *P, const float *rays, int64_t n_pixels, int32_t spp, const IrisSampler *sampler,
                        float *L, void *record, void *encoded, void *workspace, int64_t workspace_bytes, void *stream) {
    if (!s) return fail(IRIS_ERR_INVALID, "scene is NULL");
    int rc = check_params(P, true);
    if (rc) return rc;
    if (n_pixels < 0 || spp <= 0 || !sampler) return fail(IRIS_ERR_INVALID, "bad arguments");
    if (n_pixels == 0) return IRIS_OK;
    if (!rays || !L || !workspace) return fail(IRIS_ERR_INVALID, "NULL array");
    if (sampler->U && sampler->stride < 8) return fail(IRIS_ERR_INVALID, "sampler stride < 8");
    if (workspace_bytes < iris_single_workspace_bytes(n_pixels, spp)) return fail(IRIS_ERR_WORKSPACE, "workspace too small");
    if ((reinterpret_cast<uintptr_t>(workspace) & 15) || (reinterpret_cast<uintptr_t>(record) & 15) || (reinterpret_cast<uintptr_t>(encoded) & 15))
        return fail(IRIS_ERR_INVALID, "workspace/record/encoded must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_pixels * spp;
    float4 *w0 = reinterpret_cast<float4 *>(workspace), *w1 = w0 + n, *w2 = w1 + n;
    CUDA_TRY(cudaMemsetAsync(L, 0, sizeof(float) * 3 * (size_t)n_pixels, st));
    {
        ProfScope ps(K_PRIMARY, st);
        k_primary<<<blocks_for(n), IRIS_BLOCK, 0, st>>>(view_of(s), *P, *sampler, rays, n_pixels, spp, w0, w1);
    }
    LAUNCHED();
    rc = launch_field(P, n, nullptr, nullptr, w0, w1, w2, st, reinterpret_cast<__half *>(encoded),
                      /*pair=*/0);          // primary hits: the spp samples of a pixel share sectors, paired gathers do not pay
    if (rc) return rc;
    if (g_single_impl == 0) {
        ProfScope ps(K_BOUNCE_SINGLE, st);
        if (record) k_bounce_single<true><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(view_of(s), *P, *sampler, rays, n_pixels, spp, w0, w1, w2, L, reinterpret_cast<float4 *>(record));
        else k_bounce_single<false><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(view_of(s), *P, *sampler, rays, n_pixels, spp, w0, w1, w2, L, nullptr);
        LAUNCHED();
        return IRIS_OK;
    }
    // wavefront bounce: generate rays -> persistent trace -> shade, in chunks whose queue stays small.  Record mode: 0 inference,
    // 1 emitter gradient only (no `encoded` array given: the BRDF Jacobians are neither computed nor stored), 2 full
    const int rec_mode = record ? (encoded ? 2 : 1) : 0;
    const int64_t nc_max = single_chunk_samples(n);
    unsigned long long *counters = reinterpret_cast<unsigned long long *>(w2 + n);
    float4 *ro = reinterpret_cast<float4 *>(counters + SINGLE_MAX_CHUNKS), *rd = ro + 2 * nc_max, *hit = rd + 2 * nc_max, *state = hit + 2 * nc_max;
    CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * SINGLE_MAX_CHUNKS, st));
    int chunk = 0;
    for (int64_t i0 = 0; i0 < n; i0 += nc_max, ++chunk) {
        const int64_t nc = std::min(nc_max, n - i0);
        {
            ProfScope ps(K_SINGLE_GEN, st);
            if (rec_mode == 2) k_single_gen<2><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(*P, *sampler, rays, i0, nc, spp, w0, w1, w2, ro, rd, state);
            else if (rec_mode == 1) k_single_gen<1><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(*P, *sampler, rays, i0, nc, spp, w0, w1, w2, ro, rd, state);
            else k_single_gen<0><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(*P, *sampler, rays, i0, nc, spp, w0, w1, w2, ro, rd, state);
        }
        LAUNCHED();
        {
            ProfScope ps(K_TRACE_QUEUE, st);
            const int grid = (int)std::min<int64_t>(blocks_for(2 * nc), (int64_t)g_persist_ctas * (g_sm_count > 0 ? g_sm_count : 148));
            k_trace_queue<<<grid, IRIS_BLOCK, 0, st>>>(view_of(s), ro, rd, 2 * nc, nc, hit, counters + chunk, nullptr, 0);
        }
        LAUNCHED();
        {
            ProfScope ps(K_SINGLE_SHADE, st);
            float4 *rec4 = reinterpret_cast<float4 *>(record);
            if (rec_mode == 2) k_single_shade<2><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(view_of(s), *P, i0, nc, n, spp, w0, rd, hit, state, L, rec4);
            else if (rec_mode == 1) k_single_shade<1><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(view_of(s), *P, i0, nc, n, spp, w0, rd, hit, state, L, rec4);
            else k_single_shade<0><<<blocks_for(nc), IRIS_BLOCK, 0, st>>>(view_of(s), *P, i0, nc, n, spp, w0, rd, hit, state, L, nullptr);
        }
        LAUNCHED();
    }
    return IRIS_OK;
}

int iris_single_backward(const IrisShadeParams *P, const float *dL, int64_t n_pixels, int32_t spp, const void *record, const void *encoded,
                         float *d_radiance, float *d_params, void *workspace, int64_t workspace_bytes, void *stream) {
    if (!P) return fail(IRIS_ERR_INVALID, "params is NULL");
    if (n_pixels < 0 || spp <= 0) return fail(IRIS_ERR_INVALID, "bad arguments");
    if (n_pixels == 0) return IRIS_OK;
    if (!dL || !record) return fail(IRIS_ERR_INVALID, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = P->n_emitters;
    const size_t smem = K <= IRIS_BWD_KMAX ? (size_t)3 * K * IRIS_BLOCK * 4 : 0;
    const int64_t n = n_pixels * spp;
    float *d_mat = nullptr;
    if (d_params) {
        if (reinterpret_cast<uintptr_t>(d_params) & 15) return fail(IRIS_ERR_INVALID, "d_params must be 16-byte aligned (16-byte vector reductions into the grid gradient)");
        if (!encoded) return fail(IRIS_ERR_INVALID, "the field gradient needs the `encoded` array the forward was given (a record made without it holds no BRDF Jacobians)");
        if (!P->grid_f16 || !P->mlp_f16) return fail(IRIS_ERR_INVALID, "BRDF field tables missing");
        if (!workspace || workspace_bytes < iris_single_workspace_bytes(n_pixels, spp)) return fail(IRIS_ERR_WORKSPACE, "workspace too small");
    }
    if (workspace && encoded) {   // d_mat (n,5) = J^T g is left at the start of the workspace (also without d_params, for inspection)
        if (workspace_bytes < 5 * 4 * n) return fail(IRIS_ERR_WORKSPACE, "workspace too small for d_mat");
        if (reinterpret_cast<uintptr_t>(workspace) & 15) return fail(IRIS_ERR_INVALID, "workspace must be 16-byte aligned");
        d_mat = reinterpret_cast<float *>(workspace);
    }
    if (!d_radiance && !d_mat) return IRIS_OK;
    const unsigned grid = (unsigned)std::min<int64_t>(blocks_for(n), (int64_t)g_sm_count * 8);
    {
        ProfScope ps(K_SINGLE_BACKWARD, st);
        k_single_backward<<<grid, IRIS_BLOCK, d_radiance ? smem : 0, st>>>(dL, n_pixels, spp, reinterpret_cast<const float4 *>(record), K, d_radiance, d_mat);
    }
    LAUNCHED();
    if (d_params) {
        const int64_t off = ((5 * 4 * n + 15) / 16) * 16;
        if (reinterpret_cast<uintptr_t>(encoded) & 15) return fail(IRIS_ERR_INVALID, "encoded must be 16-byte aligned");
        __half *x_saved = reinterpret_cast<__half *>(const_cast<void *>(encoded));
        return run_field_backward(P, n, nullptr, reinterpret_cast<const float4 *>(record) + 5 * n, d_mat, d_params,
                                  reinterpret_cast<unsigned char *>(workspace) + off, workspace_bytes - off, st, x_saved);
    }
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------------------ wavefront estimators
int64_t iris_wave_workspace_bytes(int64_t n_lanes) { return WAVE_STREAMS * 16 * std::max<int64_t>(n_lanes, 1) + WAVE_TAIL_BYTES; }

// The live-lane lists of one estimator call (wavefront.cuh): `cur` feeds bounce k, the shade kernel appends the survivors to `next`.
struct WaveLanes {
    int32_t *cur = nullptr, *next = nullptr;
    unsigned long long *cnt = nullptr;      // cnt[k]: lanes entering bounce k
    int k = 0;
    bool on = false;
    const int32_t *idx() const { return on ? cur : nullptr; }
    const unsigned long long *count() const { return on ? cnt + k : nullptr; }
    int32_t *idx_next() const { return on ? next : nullptr; }
    unsigned long long *count_next() const { return on ? cnt + k + 1 : nullptr; }
    void advance() { std::swap(cur, next); ++k; }
};
static WaveLanes wave_lanes(const WaveState &W) {
    WaveLanes L;
    L.on = g_wave_impl == 1 && g_wave_compact == 1;       // the fused bounce kernel (wave_impl 0) works on all lanes in place
    L.cur = W.idxA;
    L.next = W.idxB;
    L.cnt = W.cnt;
    return L;
}

static int wave_field(const IrisShadeParams *P, int64_t n, float4 *pos, float4 *m1, float4 *m2, cudaStream_t st, const unsigned long long *n_dev = nullptr) {
    return launch_field(P, n, nullptr, nullptr, pos, m1, m2, st, nullptr, 1, n_dev);
}

// one "bounce a" of the wavefront estimators: fused kernel (wave_impl 0) or generate -> ray-queue trace -> resolve (default)
#define WAVE_KIND_SWITCH(KERNEL, ...)                                                         \
    switch (kind) {                                                                           \
    case 0: KERNEL<0><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(__VA_ARGS__); break;               \
    case 1: KERNEL<1><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(__VA_ARGS__); break;               \
    case 2: KERNEL<2><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(__VA_ARGS__); break;               \
    default: KERNEL<3><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(__VA_ARGS__); break;              \
    }
static int wave_bounce_a(int kind, const IrisScene *s, const IrisShadeParams *P, const IrisSampler *smp, int col0, float level, int64_t n, WaveState W,
                         const WaveLanes &L, cudaStream_t st) {
    if (g_wave_impl == 0) {
        ProfScope ps(K_WAVE_A, st);
        WAVE_KIND_SWITCH(k_wave_bounce_a, view_of(s), *P, *smp, col0, level, n, W)
        LAUNCHED();
        return IRIS_OK;
    }
    CUDA_TRY(cudaMemsetAsync(W.counter, 0, 8, st));
    {
        ProfScope ps(K_WAVE_A, st);
        WAVE_KIND_SWITCH(k_wave_gen, *P, *smp, col0, level, n, W, L.idx(), L.count())
    }
    LAUNCHED();
    {
        ProfScope ps(K_TRACE_QUEUE, st);
        const bool shadow = kind <= 1;                       // det_*: only closest-hit rays, in the first part of the queue
        const int64_t nr = shadow ? 2 * n : n;
        const int grid = (int)std::min<int64_t>(blocks_for(nr), (int64_t)g_persist_ctas * (g_sm_count > 0 ? g_sm_count : 148));
        k_trace_queue<<<grid, IRIS_BLOCK, 0, st>>>(view_of(s), W.RO, W.RD, nr, shadow ? n : 0, W.HIT, W.counter, L.count(), shadow ? 2 : 1);
    }
    LAUNCHED();
    {
        ProfScope ps(K_WAVE_A, st);
        WAVE_KIND_SWITCH(k_wave_resolve, view_of(s), n, W, L.idx(), L.count())
    }
    LAUNCHED();
    return IRIS_OK;
}

// field at the hits + radiance / MIS / state update of one bounce; advances the live-lane list
static int wave_bounce_b(int kind, const IrisShadeParams *P, int64_t n, WaveState W, WaveLanes &L, cudaStream_t st) {
    int rc = wave_field(P, n, W.H0, W.M1, W.M2, st, L.count());
    if (rc) return rc;
    {
        ProfScope ps(K_WAVE_B, st);
        if (kind == 0) k_wave_bounce_b<0><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(*P, 0.6f, 1, n, W, L.idx(), L.count(), L.idx_next(), L.count_next());
        else if (kind == 1) k_wave_bounce_b<1><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(*P, 0.6f, 1, n, W, L.idx(), L.count(), L.idx_next(), L.count_next());
        else k_wave_bounce_b<2><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(*P, 0.6f, 1, n, W, L.idx(), L.count(), L.idx_next(), L.count_next());
    }
    LAUNCHED();
    if (L.on) L.advance();
    return IRIS_OK;
}

// indirect depths (trace_indirect loop body) on the current state
static int wave_indirect(const IrisScene *s, const IrisShadeParams *P, const IrisSampler *smp, int64_t n, WaveState W, WaveLanes &L, int depth, int col_base,
                         cudaStream_t st) {
    for (int k = 0; k < depth; ++k) {
        int rc = wave_bounce_a(1, s, P, smp, col_base + 6 * k, 0.f, n, W, L, st);
        if (rc) return rc;
        if ((rc = wave_bounce_b(1, P, n, W, L, st))) return rc;
    }
    return IRIS_OK;
}

static int wave_check(const IrisScene *s, const IrisShadeParams *P, const IrisSampler *smp, int64_t n_rows, int32_t spp, int32_t depth, int need_cols,
                      void *ws, int64_t ws_bytes) {
    if (!s) return fail(IRIS_ERR_INVALID, "scene is NULL");
    int rc = check_params(P, true);
    if (rc) return rc;
    if (n_rows < 0 || spp <= 0 || depth < 0 || !smp) return fail(IRIS_ERR_INVALID, "bad arguments");
    if (depth + 2 > WAVE_MAX_BOUNCES) return fail(IRIS_ERR_INVALID, "indir_depth too large");
    if (n_rows * (int64_t)spp > 0x7FFFFFFFll) return fail(IRIS_ERR_INVALID, "too many lanes for one call (rows * spp must fit 31 bits)");
    if (smp->U && smp->stride < need_cols) return fail(IRIS_ERR_INVALID, "sampler stride too small for this estimator / depth");
    if (n_rows > 0 && (!ws || ws_bytes < iris_wave_workspace_bytes(n_rows * spp))) return fail(IRIS_ERR_WORKSPACE, "workspace too small");
    if (reinterpret_cast<uintptr_t>(ws) & 15) return fail(IRIS_ERR_INVALID, "workspace must be 16-byte aligned");
    return IRIS_OK;
}

int iris_path_tracing(const IrisScene *s, const IrisShadeParams *P, const float *rays, int64_t n_pixels, int32_t spp, int32_t indir_depth,
                      const IrisSampler *smp, float *L, void *workspace, int64_t workspace_bytes, void *stream) {
    int rc = wave_check(s, P, smp, n_pixels, spp, indir_depth, 8 + 6 * indir_depth, workspace, workspace_bytes);
    if (rc) return rc;
    if (n_pixels == 0) return IRIS_OK;
    if (!rays || !L) return fail(IRIS_ERR_INVALID, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_pixels * spp;
    WaveState W = wave_carve(workspace, n);
    WaveLanes lanes = wave_lanes(W);
    CUDA_TRY(cudaMemsetAsync(L, 0, sizeof(float) * 3 * (size_t)n_pixels, st));
    CUDA_TRY(cudaMemsetAsync(W.counter, 0, WAVE_TAIL_BYTES, st));
    {
        ProfScope ps(K_WAVE_INIT, st);
        k_wave_init_camera<<<blocks_for(n), IRIS_BLOCK, 0, st>>>(view_of(s), *P, *smp, rays, n_pixels, spp, W, lanes.on ? lanes.cur : nullptr, lanes.cnt);
    }
    LAUNCHED();
    if ((rc = wave_field(P, n, W.S0, W.S1, W.S2, st))) return rc;
    if ((rc = wave_bounce_a(0, s, P, smp, 2, 0.f, n, W, lanes, st))) return rc;
    if ((rc = wave_bounce_b(0, P, n, W, lanes, st))) return rc;
    if ((rc = wave_indirect(s, P, smp, n, W, lanes, indir_depth, 8, st))) return rc;
    {
        ProfScope ps(K_WAVE_FINISH, st);
        k_wave_finish<0><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(n_pixels, spp, W, L, nullptr);
    }
    LAUNCHED();
    return IRIS_OK;
}

int iris_path_tracing_det(const IrisScene *s, const IrisShadeParams *P, int mode, float roughness_level, const float *positions, const float *wis,
                          const float *normals, const int32_t *prim, int64_t n_pixels, int32_t spp, int32_t indir_depth, const IrisSampler *smp,
                          float *L0, float *L1, void *workspace, int64_t workspace_bytes, void *stream) {
    int rc = wave_check(s, P, smp, n_pixels, spp, indir_depth, 2 + 6 * indir_depth, workspace, workspace_bytes);
    if (rc) return rc;
    if (mode != 0 && mode != 1) return fail(IRIS_ERR_INVALID, "mode must be 0 (diffuse) or 1 (specular)");
    if (n_pixels == 0) return IRIS_OK;
    if (!positions || !wis || !normals || !prim || !L0 || (mode == 1 && !L1)) return fail(IRIS_ERR_INVALID, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_pixels * spp;
    WaveState W = wave_carve(workspace, n);
    WaveLanes lanes = wave_lanes(W);
    CUDA_TRY(cudaMemsetAsync(L0, 0, sizeof(float) * 3 * (size_t)n_pixels, st));
    if (mode == 1) CUDA_TRY(cudaMemsetAsync(L1, 0, sizeof(float) * 3 * (size_t)n_pixels, st));
    CUDA_TRY(cudaMemsetAsync(W.counter, 0, WAVE_TAIL_BYTES, st));
    {
        ProfScope ps(K_WAVE_INIT, st);
        k_wave_init_points<<<blocks_for(n), IRIS_BLOCK, 0, st>>>(positions, wis, 1, normals, prim, n_pixels, spp, W, lanes.on ? lanes.cur : nullptr, lanes.cnt);
    }
    LAUNCHED();
    if ((rc = wave_bounce_a(mode == 0 ? 2 : 3, s, P, smp, 0, mode == 0 ? 0.f : roughness_level, n, W, lanes, st))) return rc;
    if ((rc = wave_bounce_b(2, P, n, W, lanes, st))) return rc;
    if ((rc = wave_indirect(s, P, smp, n, W, lanes, indir_depth, 2, st))) return rc;
    {
        ProfScope ps(K_WAVE_FINISH, st);
        k_wave_finish<1><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(n_pixels, spp, W, L0, mode == 1 ? L1 : nullptr);
    }
    LAUNCHED();
    return IRIS_OK;
}

int iris_trace_indirect(const IrisScene *s, const IrisShadeParams *P, const float *position, const float *wo, const float *normal, int64_t n,
                        int32_t indir_depth, const IrisSampler *smp, float *L, void *workspace, int64_t workspace_bytes, void *stream) {
    int rc = wave_check(s, P, smp, n, 1, indir_depth, 6 * indir_depth, workspace, workspace_bytes);
    if (rc) return rc;
    if (n == 0) return IRIS_OK;
    if (!position || !wo || !normal || !L) return fail(IRIS_ERR_INVALID, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    WaveState W = wave_carve(workspace, n);
    WaveLanes lanes = wave_lanes(W);
    CUDA_TRY(cudaMemsetAsync(W.counter, 0, WAVE_TAIL_BYTES, st));
    {
        ProfScope ps(K_WAVE_INIT, st);
        k_wave_init_points<<<blocks_for(n), IRIS_BLOCK, 0, st>>>(position, wo, 0, normal, nullptr, n, 1, W, lanes.on ? lanes.cur : nullptr, lanes.cnt);
    }
    LAUNCHED();
    if ((rc = wave_field(P, n, W.S0, W.S1, W.S2, st))) return rc;
    if ((rc = wave_indirect(s, P, smp, n, W, lanes, indir_depth, 0, st))) return rc;
    {
        ProfScope ps(K_WAVE_FINISH, st);
        k_wave_finish<2><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(n, 1, W, L, nullptr);
    }
    LAUNCHED();
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------------------ BSDF samplers (model/brdf.py:78-210)
int iris_bsdf_sample(int mode, const float *u, int32_t u_stride, const float *wo, const float *normal, const float *mat, float roughness, int64_t n,
                     float *wi, float *pdf, float *w0, float *w1, void *stream) {
    if (mode < 0 || mode > 2) return fail(IRIS_ERR_INVALID, "mode must be 0 (sample_diffuse), 1 (sample_specular) or 2 (sample_brdf)");
    if (n < 0 || u_stride < (mode == 2 ? 3 : 2)) return fail(IRIS_ERR_INVALID, "bad bsdf_sample arguments");
    if (n == 0) return IRIS_OK;
    if (!u || !normal || !wi || (mode >= 1 && !wo) || (mode == 2 && !mat)) return fail(IRIS_ERR_INVALID, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) k_bsdf_sample<0><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(u, u_stride, wo, normal, mat, roughness, n, wi, pdf, w0, w1);
    else if (mode == 1) k_bsdf_sample<1><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(u, u_stride, wo, normal, mat, roughness, n, wi, pdf, w0, w1);
    else k_bsdf_sample<2><<<blocks_for(n), IRIS_BLOCK, 0, st>>>(u, u_stride, wo, normal, mat, roughness, n, wi, pdf, w0, w1);
    LAUNCHED();
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------------------ EmorCRF
// ------------------------------------------------------------------------------------------------ SLF bake (slf_bake.py, model/slf.py)
int iris_slf_bounds(const float *positions, const uint8_t *valid, int64_t n, int32_t reset, float *state, void *stream) {
    if (n < 0 || !state) return fail(IRIS_ERR_INVALID, "bad slf_bounds arguments");
    if (n > 0 && !positions) return fail(IRIS_ERR_INVALID, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *keys = reinterpret_cast<uint32_t *>(state) + 2;
    if (reset) {
        static const uint32_t init[2] = {0xFFFFFFFFu, 0u};
        CUDA_TRY(cudaMemcpyAsync(keys, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    if (n > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
        k_slf_bounds<<<grid, 256, 0, st>>>(positions, valid, n, keys);
        LAUNCHED();
    }
    k_slf_bounds_decode<<<1, 1, 0, st>>>(keys, state);
    LAUNCHED();
    return IRIS_OK;
}

static int slf_grid_ok(float vrange, int32_t H) {
    if (!(vrange > 0.f) || H < 1 || H > 1024) return fail(IRIS_ERR_INVALID, "bad SLF grid (range must be > 0, 1 <= H <= 1024)");
    return IRIS_OK;
}

int iris_slf_mark(const float *positions, const uint8_t *valid, int64_t n, float voxel_min, float voxel_range, int32_t H, int32_t *occupancy,
                  void *stream) {
    int rc = slf_grid_ok(voxel_range, H);
    if (rc) return rc;
    if (n < 0) return fail(IRIS_ERR_INVALID, "n < 0");
    if (n == 0) return IRIS_OK;
    if (!positions || !occupancy) return fail(IRIS_ERR_INVALID, "NULL array");
    k_slf_mark<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(positions, valid, n, voxel_min, voxel_range, H, occupancy);
    LAUNCHED();
    return IRIS_OK;
}

int64_t iris_slf_index_workspace_bytes(int32_t H) {
    const int64_t cells = (int64_t)H * H * H;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)cells);
    return 4 * cells + ((int64_t)tmp + 255) / 256 * 256 + 256;
}

int iris_slf_index(const int32_t *occupancy, int32_t H, int32_t *inds, int64_t *n_cells, void *workspace, int64_t workspace_bytes, void *stream) {
    if (H < 1 || H > 1024) return fail(IRIS_ERR_INVALID, "bad SLF grid");
    if (!occupancy || !inds || !n_cells || !workspace) return fail(IRIS_ERR_INVALID, "NULL array");
    if (workspace_bytes < iris_slf_index_workspace_bytes(H)) return fail(IRIS_ERR_WORKSPACE, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cells = (int64_t)H * H * H;
    int32_t *rank = reinterpret_cast<int32_t *>(workspace);
    void *tmp = reinterpret_cast<unsigned char *>(workspace) + ((4 * cells + 255) / 256) * 256;
    size_t tmp_bytes = (size_t)(workspace_bytes - ((4 * cells + 255) / 256) * 256);
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, occupancy, rank, (int)cells, st));
    g_launches.fetch_add(1);
    k_slf_index<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(occupancy, rank, cells, inds);
    LAUNCHED();
    int32_t last_rank = 0, last_occ = 0;
    CUDA_TRY(cudaMemcpyAsync(&last_rank, rank + cells - 1, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&last_occ, occupancy + cells - 1, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *n_cells = (int64_t)last_rank + (last_occ ? 1 : 0);
    return IRIS_OK;
}

int iris_slf_accumulate(const float *positions, const uint8_t *valid, const float *radiance, int64_t n, float voxel_min, float voxel_range, int32_t H,
                        const int32_t *inds, float *sum, int32_t *count, void *stream) {
    int rc = slf_grid_ok(voxel_range, H);
    if (rc) return rc;
    if (n < 0) return fail(IRIS_ERR_INVALID, "n < 0");
    if (n == 0) return IRIS_OK;
    if (!positions || !radiance || !inds || !sum || !count) return fail(IRIS_ERR_INVALID, "NULL array");
    k_slf_accumulate<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(positions, valid, radiance, n, voxel_min, voxel_range, H, inds, sum, count);
    LAUNCHED();
    return IRIS_OK;
}

int iris_slf_finalize(float *sum, const int32_t *count, int64_t n_cells, void *stream) {
    if (n_cells < 0) return fail(IRIS_ERR_INVALID, "n_cells < 0");
    if (n_cells == 0) return IRIS_OK;
    if (!sum || !count) return fail(IRIS_ERR_INVALID, "NULL array");
    k_slf_finalize<<<(unsigned)((3 * n_cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sum, count, n_cells);
    LAUNCHED();
    return IRIS_OK;
}

// ------------------------------------------------------------------------------------------------ emitter extraction (extract_emitter_ldr.py)
int iris_tri_accumulate(const int32_t *prim, const uint8_t *valid, const float *radiance, int64_t n, int64_t n_faces, float *tri_sum, int32_t *tri_count,
                        void *stream) {
    if (n < 0 || n_faces < 0) return fail(IRIS_ERR_INVALID, "bad tri_accumulate arguments");
    if (n == 0) return IRIS_OK;
    if (!prim || !radiance || !tri_sum || !tri_count) return fail(IRIS_ERR_INVALID, "NULL array");
    k_tri_accumulate<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prim, valid, radiance, n, n_faces, tri_sum, tri_count);
    LAUNCHED();
    return IRIS_OK;
}

int iris_emitter_classify(const float *tri_sum, const int32_t *tri_count, int64_t n_faces, float threshold, uint8_t *is_emitter, void *stream) {
    if (n_faces < 0) return fail(IRIS_ERR_INVALID, "n_faces < 0");
    if (n_faces == 0) return IRIS_OK;
    if (!tri_sum || !tri_count || !is_emitter) return fail(IRIS_ERR_INVALID, "NULL array");
    k_emitter_classify<<<(unsigned)((n_faces + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tri_sum, tri_count, n_faces, threshold, is_emitter);
    LAUNCHED();
    return IRIS_OK;
}

int iris_emitter_geometry(const float *verts, const int32_t *faces, const int64_t *emitter_faces, int64_t n_emitters, float *out_vertices, float *out_area,
                          float *out_normal, void *stream) {
    if (n_emitters < 0) return fail(IRIS_ERR_INVALID, "n_emitters < 0");
    if (n_emitters == 0) return IRIS_OK;
    if (!verts || !faces || !emitter_faces || !out_vertices || !out_area || !out_normal) return fail(IRIS_ERR_INVALID, "NULL array");
    k_emitter_geometry<<<(unsigned)((n_emitters + 127) / 128), 128, 0, (cudaStream_t)stream>>>(verts, faces, emitter_faces, n_emitters, out_vertices, out_area, out_normal);
    LAUNCHED();
    return IRIS_OK;
}

int iris_brdf_shading_forward(const float *mat, const float *diffuse, const float *specular0, const float *specular1, int32_t n_levels, int64_t n,
                              float *L, void *stream) {
    if (n < 0 || n_levels < 2) return fail(IRIS_ERR_INVALID, "bad brdf_shading arguments");
    if (n == 0) return IRIS_OK;
    if (!mat || !diffuse || !specular0 || !specular1 || !L) return fail(IRIS_ERR_INVALID, "NULL array");
    k_brdf_shading_forward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mat, diffuse, specular0, specular1, n_levels, n, L);
    LAUNCHED();
    return IRIS_OK;
}

int iris_brdf_shading_backward(const float *mat, const float *diffuse, const float *specular0, const float *specular1, int32_t n_levels, int64_t n,
                               const float *dL, float *d_mat, void *stream) {
    if (n < 0 || n_levels < 2) return fail(IRIS_ERR_INVALID, "bad brdf_shading arguments");
    if (n == 0) return IRIS_OK;
    if (!mat || !diffuse || !specular0 || !specular1 || !dL || !d_mat) return fail(IRIS_ERR_INVALID, "NULL array");
    k_brdf_shading_backward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mat, diffuse, specular0, specular1, n_levels, n, dL, d_mat);
    LAUNCHED();
    return IRIS_OK;
}

int64_t iris_denoise_workspace_bytes(int32_t height, int32_t width) { return (int64_t)std::max(height, 0) * std::max(width, 0) * 3 * (int64_t)sizeof(float); }

int iris_denoise_atrous(const float *image, const float *normal, const float *position, int32_t height, int32_t width, int32_t iterations,
                        float sigma_c, float sigma_n, float sigma_x, float *out, void *workspace, int64_t workspace_bytes, void *stream) {
    if (height < 0 || width < 0 || iterations < 0 || iterations > 12 || !(sigma_c > 0.f) || !(sigma_n >= 0.f) || !(sigma_x > 0.f))
        return fail(IRIS_ERR_INVALID, "bad denoise arguments");
    const int64_t n = (int64_t)height * width;
    if (n == 0) return IRIS_OK;
    if (!image || !out) return fail(IRIS_ERR_INVALID, "NULL array");
    if (image == out) return fail(IRIS_ERR_INVALID, "denoise: out must not alias image");
    cudaStream_t st = (cudaStream_t)stream;
    if (iterations == 0) {
        CUDA_TRY(cudaMemcpyAsync(out, image, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice, st));
        return IRIS_OK;
    }
    if (iterations > 1 && (!workspace || workspace_bytes < iris_denoise_workspace_bytes(height, width))) return fail(IRIS_ERR_WORKSPACE, "denoise: workspace too small");
    const dim3 grid((unsigned)((width + 31) / 32), (unsigned)((height + 7) / 8));
    float *tmp = reinterpret_cast<float *>(workspace);
    const float *src = image;
    for (int i = 0; i < iterations; ++i) {
        // ping-pong so that the LAST level lands in `out`
        float *dst = ((iterations - 1 - i) & 1) ? tmp : out;
        const float sc = sigma_c * ldexpf(1.0f, -i);
        k_denoise_atrous<<<grid, 256, 0, st>>>(src, normal, position, height, width, 1 << i, 1.0f / (sc * sc), sigma_n, 1.0f / (sigma_x * sigma_x), dst);
        LAUNCHED();
        src = dst;
    }
    return IRIS_OK;
}

int iris_crf_forward(const float *hdr, const float *exposure, int32_t exposure_stride, const float *crf, int32_t n_bins, int64_t n, float *ldr,
                     void *stream) {
    if (n < 0 || n_bins < 2 || (exposure_stride != 0 && exposure_stride != 1)) return fail(IRIS_ERR_INVALID, "bad crf arguments");
    if (n == 0) return IRIS_OK;
    if (!hdr || !exposure || !crf || !ldr) return fail(IRIS_ERR_INVALID, "NULL array");
    k_crf_forward<<<(unsigned)((3 * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hdr, exposure, exposure_stride, crf, n_bins, n, ldr);
    LAUNCHED();
    return IRIS_OK;
}

int iris_crf_backward(const float *hdr, const float *exposure, int32_t exposure_stride, const float *crf, int32_t n_bins, const float *d_ldr,
                      int64_t n, float *d_hdr, float *d_crf, void *stream) {
    if (n < 0 || n_bins < 2 || n_bins > 4096 || (exposure_stride != 0 && exposure_stride != 1)) return fail(IRIS_ERR_INVALID, "bad crf arguments");
    if (n == 0) return IRIS_OK;
    if (!hdr || !exposure || !crf || !d_ldr) return fail(IRIS_ERR_INVALID, "NULL array");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int rc = ensure_device_setup(dev);
    if (rc) return rc;
    const size_t smem = sizeof(float) * 3 * (size_t)n_bins;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_crf_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>((3 * n + 255) / 256, (int64_t)g_sm_count * 4);
    k_crf_backward<<<grid, 256, smem, (cudaStream_t)stream>>>(hdr, exposure, exposure_stride, crf, n_bins, d_ldr, n, d_hdr, d_crf);
    LAUNCHED();
    return IRIS_OK;
}

}  // extern "C"
