// bvh_device.cuh -- on-device builder (builder = 1): Morton-ordered LBVH (Karras 2012) -> SAH treelets -> SAH-optimal collapse to 8-wide ->
// octant slot assignment -> conservative 8-bit quantisation, producing exactly the node / triangle layout of bvh8.h.
// Milliseconds instead of the host builder's ~0.7 s per million triangles.  The lower levels (subtrees of <= IRIS_SAH_TREELET
// primitives) are rebuilt with a binned SAH in shared memory (k_lbvh_sah_treelets); the levels above them are the Morton splits.
// Hit results do not depend on the builder: traversal is exact (tests run both).
#pragma once
#include <cub/cub.cuh>

#include "bvh8.h"
#include "common.cuh"

struct DBox {
    float lo[3], hi[3];
};

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {   // spread 21 bits to every third bit
    v &= 0x1FFFFFull;
    v = (v | v << 32) & 0x1F00000000FFFFull;
    v = (v | v << 16) & 0x1F0000FF0000FFull;
    v = (v | v << 8) & 0x100F00F00F00F00Full;
    v = (v | v << 4) & 0x10C30C30C30C30C3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__device__ __forceinline__ void atomic_min_f(float *a, float v) {
    int *ai = reinterpret_cast<int *>(a);
    int old = *ai;
    while (v < __int_as_float(old)) {
        const int prev = atomicCAS(ai, old, __float_as_int(v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
    int *ai = reinterpret_cast<int *>(a);
    int old = *ai;
    while (v > __int_as_float(old)) {
        const int prev = atomicCAS(ai, old, __float_as_int(v));
        if (prev == old) break;
        old = prev;
    }
}

// records + unpadded boxes + scene bounds (bounds[0..2] = lo, [3..5] = hi)
__global__ void k_lbvh_prepare(const float *__restrict__ verts, const int32_t *__restrict__ faces, int n, TriRecord *recs, DBox *tbox, float *bounds) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (f < n) {
        const float *a = verts + 3 * (int64_t)faces[3 * f], *b = verts + 3 * (int64_t)faces[3 * f + 1], *c = verts + 3 * (int64_t)faces[3 * f + 2];
        TriRecord r;
        for (int k = 0; k < 3; ++k) {
            r.v0[k] = a[k];
            r.e1[k] = __fsub_rn(b[k], a[k]);
            r.e2[k] = __fsub_rn(c[k], a[k]);
            const float p1 = __fadd_rn(r.v0[k], r.e1[k]), p2 = __fadd_rn(r.v0[k], r.e2[k]);
            lo[k] = fminf(fminf(a[k], fminf(b[k], c[k])), fminf(p1, p2));
            hi[k] = fmaxf(fmaxf(a[k], fmaxf(b[k], c[k])), fmaxf(p1, p2));
        }
        const float cx = __fsub_rn(__fmul_rn(r.e1[1], r.e2[2]), __fmul_rn(r.e1[2], r.e2[1]));
        const float cy = __fsub_rn(__fmul_rn(r.e1[2], r.e2[0]), __fmul_rn(r.e1[0], r.e2[2]));
        const float cz = __fsub_rn(__fmul_rn(r.e1[0], r.e2[1]), __fmul_rn(r.e1[1], r.e2[0]));
        if (cx == 0.f && cy == 0.f && cz == 0.f)
            for (int k = 0; k < 3; ++k) r.e1[k] = r.e2[k] = 0.f;      // zero-area triangle: never hit (oracle rule)
        r.prim = f;
        r.pad[0] = r.pad[1] = 0;
        recs[f] = r;
        DBox bx;
        for (int k = 0; k < 3; ++k) { bx.lo[k] = lo[k]; bx.hi[k] = hi[k]; }
        tbox[f] = bx;
    }
    // block reduction of the bounds, one atomic per block and component
    __shared__ float s[6][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 3; ++k) {
        float a = lo[k], b = hi[k];
        for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
        if (lane == 0) { s[k][warp] = a; s[3 + k][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[threadIdx.x][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = threadIdx.x < 3 ? fminf(v, s[threadIdx.x][w]) : fmaxf(v, s[threadIdx.x][w]);
        if (threadIdx.x < 3) atomic_min_f(bounds + threadIdx.x, v); else atomic_max_f(bounds + threadIdx.x, v);
    }
}

__global__ void k_lbvh_morton(const DBox *__restrict__ tbox, int n, const float *__restrict__ bounds, unsigned long long *keys, uint32_t *idx) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    unsigned long long code = 0;
    for (int k = 0; k < 3; ++k) {
        const float c = 0.5f * (tbox[f].lo[k] + tbox[f].hi[k]);
        const float ext = fmaxf(bounds[3 + k] - bounds[k], 1e-30f);
        const float u = fminf(fmaxf((c - bounds[k]) / ext, 0.f), 1.f);
        const unsigned long long q = (unsigned long long)fminf(u * 2097152.f, 2097151.f);
        code |= expand21(q) << (2 - k);
    }
    keys[f] = code;
    idx[f] = (uint32_t)f;
}

// Binary radix tree.  Nodes: internal i in [0,n-1), leaf k at n-1+k.
struct LbvhNodes {
    int2 *child;      // internal: (left, right) node ids
    int *parent;      // all 2n-1 nodes
    int2 *range;      // internal: [first,last] sorted leaf range
    DBox *box;        // all 2n-1 nodes (padded)
    int *flag;        // internal: arrival counter
    float *cost;      // internal: 8 per node, cost[i] = C(node, i) for i = 1..7 (see k_lbvh_fit)
    uint32_t *dec;    // internal: the decisions behind cost[]
};

__device__ __forceinline__ int lbvh_delta(const unsigned long long *__restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long x = keys[i] ^ keys[j];
    return x == 0ull ? 64 + __clz(i ^ j) : __clzll((long long)x);
}

__global__ void k_lbvh_hierarchy(const unsigned long long *__restrict__ keys, int n, LbvhNodes N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = first == gamma ? (n - 1 + gamma) : gamma;
    const int right = last == gamma + 1 ? (n - 1 + gamma + 1) : gamma + 1;
    N.child[i] = make_int2(left, right);
    N.range[i] = make_int2(first, last);
    N.parent[left] = i;
    N.parent[right] = i;
    if (i == 0) N.parent[0] = -1;
}

// ------------------------------------------------------------------------------------------------------------------------
// SAH treelets.  The Morton hierarchy is good at the top (spatial median splits of the whole scene) and poor at the bottom, where
// neighbouring triangles of different size and shape are grouped by centroid code alone (on a scan-like mesh the host binned-SAH tree
// traces ~10 % faster, profiles/r2l_*).  Every subtree of at most IRIS_SAH_TREELET primitives whose parent is larger is therefore
// REBUILT by one CTA with a binned SAH (SAH_BINS bins x 3 axes over the centroids, cost = area x count) entirely in shared memory: the
// subtree's primitives are re-ordered inside their range of the sorted array and its internal nodes re-linked, using the same
// numbering rule as the radix tree (split after sorted position g: an internal left child is node g, an internal right child node
// g + 1), so the ids stay unique, the subtree's root keeps its id, and the fit / collapse passes below see an ordinary tree.
// ------------------------------------------------------------------------------------------------------------------------
#ifndef IRIS_SAH_TREELET
#define IRIS_SAH_TREELET 4096     // primitives per rebuilt subtree (44 B of shared memory each); 0: plain LBVH
#endif
#ifndef SAH_BINS
#define SAH_BINS 16
#endif
#define SAH_THREADS 128
#define SAH_SMEM_BYTES(T) ((size_t)(T) * 44)

__device__ __forceinline__ int sah_f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float sah_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void k_lbvh_treelet_roots(int n, LbvhNodes N, int T, int *list, int *count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 r = N.range[i];
    const int c = r.y - r.x + 1;
    if (c > T || c < 4) return;
    const int p = N.parent[i];
    if (p >= 0) {
        const int2 pr = N.range[p];
        if (pr.y - pr.x + 1 <= T) return;
    }
    list[atomicAdd(count, 1)] = i;
}

// ---- optional (iris_set_option lbvh_sah_top 1; off by default: measured slower on whole interior views)
// ---- top of the tree: the maximal subtrees of <= T primitives ("clusters", single leaves included) tile the sorted array; their boxes
//      go to the host, which builds a SAH tree over the few hundred / thousand of them (host_sah_top) and sends back the new cluster
//      order and the top nodes; the clusters are moved to their new places and the top nodes re-linked with the radix-tree numbering.
__global__ void k_lbvh_clusters(int n, LbvhNodes N, int T, int2 *out, int *count) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= 2 * n - 1) return;
    int start, cnt;
    if (node >= n - 1) { start = node - (n - 1); cnt = 1; } else { const int2 r = N.range[node]; start = r.x; cnt = r.y - r.x + 1; }
    if (cnt > T) return;
    const int p = N.parent[node];
    if (p >= 0) {
        const int2 pr = N.range[p];
        if (pr.y - pr.x + 1 <= T) return;
    }
    out[atomicAdd(count, 1)] = make_int2(start, cnt);
}
__global__ void k_lbvh_cluster_boxes(const int2 *__restrict__ cl, int C, const DBox *__restrict__ tbox, const uint32_t *__restrict__ sorted, float *cbox) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = lane; k < cl[c].y; k += 32) {
        const DBox b = tbox[sorted[cl[c].x + k]];
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], b.lo[a]); hi[a] = fmaxf(hi[a], b.hi[a]); }
    }
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    if (lane == 0)
        for (int a = 0; a < 3; ++a) { cbox[6 * c + a] = lo[a]; cbox[6 * c + 3 + a] = hi[a]; }
}
__global__ void k_lbvh_permute(const int3 *__restrict__ perm /* old start, size, new start */, const uint32_t *__restrict__ src, uint32_t *dst) {
    const int3 p = perm[blockIdx.x];
    for (int k = threadIdx.x; k < p.y; k += blockDim.x) dst[p.z + k] = src[p.x + k];
}
__global__ void k_lbvh_write_top(const int *__restrict__ rec /* id, left, right, first, last */, int n_rec, const int3 *__restrict__ roots, int n_roots, LbvhNodes N) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_rec) {
        const int id = rec[5 * k], l = rec[5 * k + 1], r = rec[5 * k + 2];
        N.child[id] = make_int2(l, r);
        N.range[id] = make_int2(rec[5 * k + 3], rec[5 * k + 4]);
        N.parent[l] = id;
        N.parent[r] = id;
        if (id == 0) N.parent[0] = -1;
    }
    if (k < n_roots) N.range[roots[k].x] = make_int2(roots[k].y, roots[k].z);
}

template <int T>
__global__ void __launch_bounds__(SAH_THREADS) k_lbvh_sah_treelets(const int *__restrict__ list, int n_list, int n, LbvhNodes N,
                                                                   const DBox *__restrict__ tbox, uint32_t *sorted) {
    extern __shared__ __align__(16) unsigned char sah_smem[];    // 44 bytes per primitive: T = 1024 -> 44 KB, 4096 -> 176 KB
    float (*s_c)[T] = reinterpret_cast<float (*)[T]>(sah_smem);
    float (*s_lo)[T] = s_c + 3, (*s_hi)[T] = s_c + 6;
    uint32_t *s_id = reinterpret_cast<uint32_t *>(s_c + 9);
    uint16_t *s_ord = reinterpret_cast<uint16_t *>(s_id + T), *s_ord2 = s_ord + T;
    __shared__ int s_bins[3][SAH_BINS][7];      // count | lo xyz | hi xyz (order-preserving int images of the floats)
    __shared__ int s_cb[6];
    __shared__ int s_split[3];                  // axis (-1: none), bin, primitives on the left
    __shared__ float s_cost[3 * (SAH_BINS - 1)];
    __shared__ int s_nl[3 * (SAH_BINS - 1)];
    __shared__ int s_warp[2 * (SAH_THREADS / 32)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = blockIdx.x; t < n_list; t += gridDim.x) {
        const int root = list[t];
        const int2 rr = N.range[root];
        const int a = rr.x, m = rr.y - rr.x + 1;
        __syncthreads();
        for (int k = tid; k < m; k += SAH_THREADS) {
            const uint32_t id = sorted[a + k];
            const DBox b = tbox[id];
            s_id[k] = id;
            for (int ax = 0; ax < 3; ++ax) { s_lo[ax][k] = b.lo[ax]; s_hi[ax][k] = b.hi[ax]; s_c[ax][k] = 0.5f * (b.lo[ax] + b.hi[ax]); }
            s_ord[k] = (uint16_t)k;
        }
        __syncthreads();
        // iterative top-down build; (f, l, id) and the stack are block-uniform
        int3 stack[24];
        int sp = 0;
        int f = 0, l = m - 1, id = root;
        for (;;) {
            const int c = l - f + 1;
            int n_l = c >> 1;                                       // default: split the current order in the middle
            if (c > 3) {
                if (tid < 6) s_cb[tid] = tid < 3 ? 0x7FFFFFFF : (int)0x80000000;
                for (int k = tid; k < 3 * SAH_BINS * 7; k += SAH_THREADS) {
                    const int w = k % 7;
                    (&s_bins[0][0][0])[k] = w == 0 ? 0 : (w < 4 ? 0x7FFFFFFF : (int)0x80000000);
                }
                __syncthreads();
                for (int k = f + tid; k <= l; k += SAH_THREADS) {
                    const int j = s_ord[k];
                    for (int ax = 0; ax < 3; ++ax) {
                        const int o = sah_f2ord(s_c[ax][j]);
                        atomicMin(&s_cb[ax], o);
                        atomicMax(&s_cb[3 + ax], o);
                    }
                }
                __syncthreads();
                float c0[3], sc[3];
                for (int ax = 0; ax < 3; ++ax) {
                    c0[ax] = sah_ord2f(s_cb[ax]);
                    const float ext = sah_ord2f(s_cb[3 + ax]) - c0[ax];
                    sc[ax] = ext > 0.f ? (float)SAH_BINS / ext : 0.f;
                }
                for (int k = f + tid; k <= l; k += SAH_THREADS) {
                    const int j = s_ord[k];
                    for (int ax = 0; ax < 3; ++ax) {
                        if (!(sc[ax] > 0.f)) continue;
                        const int b = min(SAH_BINS - 1, max(0, (int)((s_c[ax][j] - c0[ax]) * sc[ax])));
                        int *B = s_bins[ax][b];
                        atomicAdd(B, 1);
                        for (int d = 0; d < 3; ++d) {
                            atomicMin(B + 1 + d, sah_f2ord(s_lo[d][j]));
                            atomicMax(B + 4 + d, sah_f2ord(s_hi[d][j]));
                        }
                    }
                }
                __syncthreads();
                // one thread per candidate (axis, last bin on the left), then the minimum over the candidates by warp 0
                if (tid < 3 * (SAH_BINS - 1)) {
                    float cost = INFINITY;
                    int nl = 0;
                    const int ax = tid / (SAH_BINS - 1), kb = tid % (SAH_BINS - 1);
                    if (sc[ax] > 0.f) {
                        float lo0[3] = {INFINITY, INFINITY, INFINITY}, hi0[3] = {-INFINITY, -INFINITY, -INFINITY};
                        float lo1[3] = {INFINITY, INFINITY, INFINITY}, hi1[3] = {-INFINITY, -INFINITY, -INFINITY};
                        int n0 = 0, n1 = 0;
                        for (int b = 0; b < SAH_BINS; ++b) {
                            const int *B = s_bins[ax][b];
                            if (B[0] == 0) continue;
                            if (b <= kb) {
                                n0 += B[0];
                                for (int d = 0; d < 3; ++d) { lo0[d] = fminf(lo0[d], sah_ord2f(B[1 + d])); hi0[d] = fmaxf(hi0[d], sah_ord2f(B[4 + d])); }
                            } else {
                                n1 += B[0];
                                for (int d = 0; d < 3; ++d) { lo1[d] = fminf(lo1[d], sah_ord2f(B[1 + d])); hi1[d] = fmaxf(hi1[d], sah_ord2f(B[4 + d])); }
                            }
                        }
                        if (n0 > 0 && n1 > 0) {
                            const float e0x = hi0[0] - lo0[0], e0y = hi0[1] - lo0[1], e0z = hi0[2] - lo0[2];
                            const float e1x = hi1[0] - lo1[0], e1y = hi1[1] - lo1[1], e1z = hi1[2] - lo1[2];
                            cost = (e0x * e0y + e0y * e0z + e0z * e0x) * (float)n0 + (e1x * e1y + e1y * e1z + e1z * e1x) * (float)n1;
                            nl = n0;
                        }
                    }
                    s_cost[tid] = cost;
                    s_nl[tid] = nl;
                }
                __syncthreads();
                if (warp == 0) {
                    float cost = INFINITY;
                    int best = 0;
                    for (int k = lane; k < 3 * (SAH_BINS - 1); k += 32)
                        if (s_cost[k] < cost) { cost = s_cost[k]; best = k; }
                    for (int o = 16; o > 0; o >>= 1) {
                        const float oc = __shfl_xor_sync(0xffffffffu, cost, o);
                        const int ob = __shfl_xor_sync(0xffffffffu, best, o);
                        if (oc < cost || (oc == cost && ob < best)) { cost = oc; best = ob; }
                    }
                    if (lane == 0) {
                        const bool ok = cost < INFINITY;
                        s_split[0] = ok ? best / (SAH_BINS - 1) : -1;
                        s_split[1] = best % (SAH_BINS - 1);
                        s_split[2] = s_nl[best];
                    }
                }
                __syncthreads();
                const int ax = s_split[0], kb = s_split[1];
                if (ax >= 0) {
                    n_l = s_split[2];
                    int run_l = 0, run_r = 0;
                    for (int base = f; base <= l; base += SAH_THREADS) {
                        const int k = base + tid;
                        const bool in = k <= l;
                        int j = 0;
                        bool left = false;
                        if (in) {
                            j = s_ord[k];
                            left = min(SAH_BINS - 1, max(0, (int)((s_c[ax][j] - c0[ax]) * sc[ax]))) <= kb;
                        }
                        const unsigned bl = __ballot_sync(0xffffffffu, in && left), br = __ballot_sync(0xffffffffu, in && !left);
                        if (lane == 0) { s_warp[warp] = __popc(bl); s_warp[SAH_THREADS / 32 + warp] = __popc(br); }
                        __syncthreads();
                        int pl = 0, pr = 0, tl = 0, tr = 0;
                        for (int w = 0; w < SAH_THREADS / 32; ++w) {
                            const int x = s_warp[w], y = s_warp[SAH_THREADS / 32 + w];
                            if (w < warp) { pl += x; pr += y; }
                            tl += x; tr += y;
                        }
                        const unsigned lt = (1u << lane) - 1u;
                        if (in) {
                            const int pos = left ? f + run_l + pl + __popc(bl & lt) : f + n_l + run_r + pr + __popc(br & lt);
                            s_ord2[pos] = (uint16_t)j;
                        }
                        run_l += tl;
                        run_r += tr;
                        __syncthreads();
                    }
                    for (int k = f + tid; k <= l; k += SAH_THREADS) s_ord[k] = s_ord2[k];
                    __syncthreads();
                }
            }
            const int mid = f + n_l, n_r = c - n_l;
            const int left_id = n_l == 1 ? n - 1 + a + f : a + mid - 1;
            const int right_id = n_r == 1 ? n - 1 + a + mid : a + mid;
            if (tid == 0) {
                N.child[id] = make_int2(left_id, right_id);
                N.range[id] = make_int2(a + f, a + l);
                N.parent[left_id] = id;
                N.parent[right_id] = id;
            }
            // continue with the smaller side, park the larger one (stack depth <= log2 T)
            const bool go_left_first = n_l <= n_r;
            const int3 L = make_int3(f, mid - 1, left_id), R = make_int3(mid, l, right_id);
            const int3 first = go_left_first ? L : R, second = go_left_first ? R : L;
            if (second.y > second.x) stack[sp++] = second;
            if (first.y > first.x) { f = first.x; l = first.y; id = first.z; continue; }
            if (sp == 0) break;
            const int3 nx = stack[--sp];
            f = nx.x; l = nx.y; id = nx.z;
        }
        __syncthreads();
        for (int k = tid; k < m; k += SAH_THREADS) sorted[a + k] = s_id[s_ord[k]];
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// SAH-optimal collapse to 8-wide (Ylitie, Karras, Laine 2017, section 3.1), computed bottom-up next to the boxes:
//   C(n, i) = least cost of turning the subtree of binary node n into AT MOST i children of a wide node, i = 1..7
//   C(n, 1) = min( leaf: A_n P_n c_prim  (P_n <= 3 triangles),  inner wide node: A_n c_node + D(n, 8) )
//   C(n, i) = min( D(n, i), C(n, i - 1) ),   D(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
// dec packs the argmins: bits [3(j-2), 3(j-2)+3) = k for D(n, j), j = 2..8; bit 21 + (i-2) = "C(n,i) is C(n,i-1)", i = 2..7; bit 27 = leaf.
// k_lbvh_collapse then expands every wide node along these decisions instead of greedily opening the child of largest area.
// ------------------------------------------------------------------------------------------------------------------------
#ifndef IRIS_SLOT_2OPT
#define IRIS_SLOT_2OPT 1
#endif
#ifndef IRIS_COLLAPSE_DP
#define IRIS_COLLAPSE_DP 1
#endif
#ifndef IRIS_SAH_CPRIM
#define IRIS_SAH_CPRIM 0.6f       // cost of one triangle test relative to one 8-wide node visit
#endif
__device__ __forceinline__ float dbox_area(const DBox &b);
__device__ __forceinline__ void lbvh_child_costs(const LbvhNodes &N, int n, int node, float area, float c[8]) {
    if (node >= n - 1) {                                        // a single triangle: one leaf slot whatever i is
#pragma unroll
        for (int i = 1; i < 8; ++i) c[i] = area * IRIS_SAH_CPRIM;
    } else {
#pragma unroll
        for (int i = 1; i < 8; ++i) c[i] = __ldcg(N.cost + 8 * (size_t)node + i);      // (written by another thread of this launch: read at L2)
    }
}
__device__ __forceinline__ void lbvh_collapse_cost(const LbvhNodes &N, int n, int p, int2 ch, float area_l, float area_r, float area_p) {
    float cl[8], cr[8], D[9];
    lbvh_child_costs(N, n, ch.x, area_l, cl);
    lbvh_child_costs(N, n, ch.y, area_r, cr);
    uint32_t dec = 0;
#pragma unroll
    for (int j = 2; j <= 8; ++j) {
        float best = INFINITY;
        int bk = 1;
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            if (k >= j || j - k > 7) continue;
            const float v = cl[k] + cr[j - k];
            if (v < best) { best = v; bk = k; }
        }
        D[j] = best;
        dec |= (uint32_t)bk << (3 * (j - 2));
    }
    const int2 r = N.range[p];
    const int cnt = r.y - r.x + 1;
    float c1 = area_p + D[8];
    if (cnt <= IRIS_MAX_LEAF) {
        const float leaf = area_p * (float)cnt * IRIS_SAH_CPRIM;
        if (leaf <= c1) { c1 = leaf; dec |= 1u << 27; }
    }
    float *C = N.cost + 8 * (size_t)p;
    C[1] = c1;
    float prev = c1;
#pragma unroll
    for (int i = 2; i < 8; ++i) {
        float v = D[i];
        if (prev <= v) { v = prev; dec |= 1u << (21 + (i - 2)); }
        C[i] = v;
        prev = v;
    }
    N.dec[p] = dec;
}

__global__ void k_lbvh_fit(const DBox *__restrict__ tbox, const uint32_t *__restrict__ sorted, int n, const float *__restrict__ bounds, LbvhNodes N) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float ext = 0.f, amax = 0.f;
    for (int a = 0; a < 3; ++a) {
        ext = fmaxf(ext, bounds[3 + a] - bounds[a]);
        amax = fmaxf(amax, fmaxf(fabsf(bounds[a]), fabsf(bounds[3 + a])));
    }
    const float pad = 1e-5f * fmaxf(ext, amax) + 1e-30f;      // same conservative padding as the host builder
    DBox b = tbox[sorted[k]];
    for (int a = 0; a < 3; ++a) { b.lo[a] -= pad; b.hi[a] += pad; }
    int node = n - 1 + k;
    N.box[node] = b;
    if (n == 1) return;
    int p = N.parent[node];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(&N.flag[p], 1) == 0) return;            // first child to arrive leaves; the second one fits the parent
        const int2 c = N.child[p];
        const DBox x = N.box[c.x], y = N.box[c.y];
        for (int a = 0; a < 3; ++a) { b.lo[a] = fminf(x.lo[a], y.lo[a]); b.hi[a] = fmaxf(x.hi[a], y.hi[a]); }
        N.box[p] = b;
#if IRIS_COLLAPSE_DP
        lbvh_collapse_cost(N, n, p, c, dbox_area(x), dbox_area(y), dbox_area(b));
#endif
        p = N.parent[p];
    }
}

__device__ __forceinline__ float dbox_area(const DBox &b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return 2.f * (dx * dy + dy * dz + dz * dx);
}
__device__ __forceinline__ int lbvh_count(const LbvhNodes &N, int n, int node) {
    if (node >= n - 1) return 1;
    const int2 r = N.range[node];
    return r.y - r.x + 1;
}
__device__ __forceinline__ int lbvh_first(const LbvhNodes &N, int n, int node) { return node >= n - 1 ? node - (n - 1) : N.range[node].x; }

__device__ __forceinline__ uint8_t dev_exp_for_extent(double ext) {
    if (!(ext > 0.0)) return (uint8_t)1;
    int e = (int)ceil(log2(ext / (double)BVH8_QMAX));
    while (ldexp((double)BVH8_QMAX, e) < ext) ++e;
    e = max(-126, min(127, e));
    return (uint8_t)(e + 127);
}

// One thread per wide node of the current level [begin,end): wide_bin[w] = binary node it represents.
__global__ void k_lbvh_collapse(int begin, int end, int n, LbvhNodes N, const TriRecord *__restrict__ recs, const uint32_t *__restrict__ sorted,
                                int *wide_bin, int *wide_depth, Bvh8Node *wide, TriRecord *tris_out, int *counters /* [0]=nodes [1]=tris [2]=max depth */) {
    const int w = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= end) return;
    const int bnode = wide_bin[w];
    int ch[8];
    int nch = 0;
#if IRIS_COLLAPSE_DP
    bool leaf_ch[8];                                         // child i is a leaf slot (<= 3 triangles) rather than a wide node of its own
    if (bnode >= n - 1 || ((N.dec[bnode] >> 27) & 1u)) {
        leaf_ch[nch] = true;
        ch[nch++] = bnode;                                   // tiny scene: the root itself is a leaf child
    } else {
        int st_node[8], st_i[8], sp = 0;
        st_node[sp] = bnode; st_i[sp] = 8; ++sp;             // distribute the root over 8 slots
        bool root = true;
        while (sp > 0) {
            --sp;
            const int nd = st_node[sp];
            int i = st_i[sp];
            if (nd >= n - 1) { leaf_ch[nch] = true; ch[nch++] = nd; continue; }       // a single triangle
            const uint32_t dec = N.dec[nd];
            if (!root) {
                while (i > 1 && ((dec >> (21 + (i - 2))) & 1u)) --i;                    // C(nd, i) = C(nd, i - 1)
                if (i == 1) { leaf_ch[nch] = (dec >> 27) & 1u; ch[nch++] = nd; continue; }
            }
            root = false;
            const int k = (int)((dec >> (3 * (i - 2))) & 7u);
            const int2 c = N.child[nd];
            st_node[sp] = c.y; st_i[sp] = i - k; ++sp;       // (right pushed first: the left subtree comes out first)
            st_node[sp] = c.x; st_i[sp] = k; ++sp;
        }
    }
#define LBVH_IS_INNER(i) (!leaf_ch[i])
#else
    if (lbvh_count(N, n, bnode) <= IRIS_MAX_LEAF) {
        ch[nch++] = bnode;                                   // tiny scene: the root itself is a leaf child
    } else {
        const int2 c = N.child[bnode];
        ch[nch++] = c.x;
        ch[nch++] = c.y;
        while (nch < 8) {
            int best = -1;
            float ba = -1.f;
            for (int i = 0; i < nch; ++i)
                if (lbvh_count(N, n, ch[i]) > IRIS_MAX_LEAF) {
                    const float a = dbox_area(N.box[ch[i]]);
                    if (a > ba) { ba = a; best = i; }
                }
            if (best < 0) break;
            const int2 c2 = N.child[ch[best]];
            ch[best] = c2.x;
            ch[nch++] = c2.y;
        }
    }
#define LBVH_IS_INNER(i) (lbvh_count(N, n, ch[i]) > IRIS_MAX_LEAF)
#endif
    DBox nb;
    for (int a = 0; a < 3; ++a) { nb.lo[a] = INFINITY; nb.hi[a] = -INFINITY; }
    for (int i = 0; i < nch; ++i) {
        const DBox b = N.box[ch[i]];
        for (int a = 0; a < 3; ++a) { nb.lo[a] = fminf(nb.lo[a], b.lo[a]); nb.hi[a] = fmaxf(nb.hi[a], b.hi[a]); }
    }
    // slot assignment: greedy minimum of dot(centroid_child - centroid_node, D_slot)
    int child_in_slot[8], slot_of[8];
    for (int s = 0; s < 8; ++s) { child_in_slot[s] = -1; slot_of[s] = -1; }
    float cost[8][8];
    for (int i = 0; i < nch; ++i) {
        const DBox b = N.box[ch[i]];
        float cc[3];
        for (int a = 0; a < 3; ++a) cc[a] = 0.5f * (b.lo[a] + b.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
        for (int s = 0; s < 8; ++s) cost[i][s] = cc[0] * ((s & 4) ? -1.f : 1.f) + cc[1] * ((s & 2) ? -1.f : 1.f) + cc[2] * ((s & 1) ? -1.f : 1.f);
    }
    for (int it = 0; it < nch; ++it) {
        float bc = INFINITY;
        int bi = -1, bs = -1;
        for (int i = 0; i < nch; ++i) {
            if (slot_of[i] >= 0) continue;
            for (int s = 0; s < 8; ++s)
                if (child_in_slot[s] < 0 && cost[i][s] < bc) { bc = cost[i][s]; bi = i; bs = s; }
        }
        slot_of[bi] = bs;
        child_in_slot[bs] = bi;
    }
#if IRIS_SLOT_2OPT
    // pairwise improvement of the greedy assignment: exchange the contents of two slots (children or holes) while the total cost drops
    for (int pass = 0; pass < 4; ++pass) {
        bool any = false;
        for (int s0 = 0; s0 < 8; ++s0)
            for (int s1 = s0 + 1; s1 < 8; ++s1) {
                const int i0 = child_in_slot[s0], i1 = child_in_slot[s1];
                if (i0 < 0 && i1 < 0) continue;
                const float cur = (i0 >= 0 ? cost[i0][s0] : 0.f) + (i1 >= 0 ? cost[i1][s1] : 0.f);
                const float alt = (i0 >= 0 ? cost[i0][s1] : 0.f) + (i1 >= 0 ? cost[i1][s0] : 0.f);
                if (alt < cur) {
                    child_in_slot[s0] = i1; child_in_slot[s1] = i0;
                    if (i0 >= 0) slot_of[i0] = s1;
                    if (i1 >= 0) slot_of[i1] = s0;
                    any = true;
                }
            }
        if (!any) break;
    }
#endif
    int n_inner = 0, n_tris = 0;
    for (int i = 0; i < nch; ++i) {
        const int c = lbvh_count(N, n, ch[i]);
        if (LBVH_IS_INNER(i)) ++n_inner; else n_tris += c;
    }
    const int child_base = n_inner ? atomicAdd(&counters[0], n_inner) : 0;
    const int tri_base = n_tris ? atomicAdd(&counters[1], n_tris) : 0;
    const int depth = wide_depth[w];
    if (n_inner) atomicMax(&counters[2], depth + 1);
    Bvh8Node O;
    memset(&O, 0, sizeof(O));
    for (int a = 0; a < 3; ++a) {
        O.p[a] = nb.lo[a];
        O.e[a] = dev_exp_for_extent((double)nb.hi[a] - (double)nb.lo[a]);
    }
    O.child_base = (uint32_t)child_base;
    O.tri_base = (uint32_t)tri_base;
    int inner_i = 0, tri_off = 0;
    for (int s = 0; s < 8; ++s) {
        const int i = child_in_slot[s];
        if (i < 0) continue;
        const DBox b = N.box[ch[i]];
        bvh8_q_t *qlo[3] = {O.qlo_x, O.qlo_y, O.qlo_z}, *qhi[3] = {O.qhi_x, O.qhi_y, O.qhi_z};
        for (int a = 0; a < 3; ++a) {
            const double sc = ldexp(1.0, (int)O.e[a] - 127);
            qlo[a][s] = bvh8_encode_q((int)fmax(0.0, fmin((double)BVH8_QMAX, floor(((double)b.lo[a] - (double)O.p[a]) / sc))));
            qhi[a][s] = bvh8_encode_q((int)fmax(0.0, fmin((double)BVH8_QMAX, ceil(((double)b.hi[a] - (double)O.p[a]) / sc))));
        }
        const int c = lbvh_count(N, n, ch[i]);
        if (LBVH_IS_INNER(i)) {
            O.imask |= (uint8_t)(1u << s);
            O.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            wide_bin[child_base + inner_i] = ch[i];
            wide_depth[child_base + inner_i] = depth + 1;
            ++inner_i;
        } else {
            const uint32_t unary = c == 1 ? 1u : (c == 2 ? 3u : 7u);
            O.meta[s] = (uint8_t)((unary << 5) | (uint32_t)tri_off);
            const int first = lbvh_first(N, n, ch[i]);
            for (int t = 0; t < c; ++t) tris_out[tri_base + tri_off + t] = recs[sorted[first + t]];
            tri_off += c;
        }
    }
    wide[w] = O;
}
