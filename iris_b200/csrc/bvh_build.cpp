// bvh_build.cpp -- host builder: binned-SAH binary BVH -> greedy collapse to 8-wide -> octant slot assignment
// -> conservative 8-bit quantisation (layout in bvh8.h).  Replaces the OptiX acceleration-structure build that
// mitsuba.load_dict triggers in the reference (train_emitter.py:57-63); B200 has no RT cores, so the structure
// is ours.  Build time is reported by iris_scene_stats and excluded from throughput numbers.
#include "bvh8.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; } }
    void grow(const Box &b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    void grow(const float *p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.f)) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Node2 {
    Box box;
    int32_t left;   // inner: index of left child (right = left + 1); leaf: first primitive
    int32_t count;  // 0 = inner
    int32_t first;  // first primitive of the subtree (primitives of a subtree are contiguous in `order`)
    int32_t total;  // primitives in the subtree
};

constexpr int kBins = 16;
constexpr int kMaxLeaf = IRIS_MAX_LEAF;

struct Builder {
    int32_t max_leaf = kMaxLeaf;   // 1: split down to single primitives (top tree over clusters, host_sah_top)
    std::vector<int32_t> weight;   // optional: primitives behind each box (clusters); empty = 1 each.  Only the SAH cost uses it.
    std::vector<Box> tbox;
    std::vector<float> cent;   // 3 per tri
    std::vector<int32_t> order;
    std::vector<Node2> nodes;

    void build_range(int32_t node, int32_t first, int32_t count) {
        // iterative with an explicit stack to keep deep scenes off the C stack
        struct Job { int32_t node, first, count; };
        std::vector<Job> st;
        st.push_back({node, first, count});
        while (!st.empty()) {
            Job j = st.back();
            st.pop_back();
            Box b, cb;
            b.reset();
            cb.reset();
            for (int32_t i = j.first; i < j.first + j.count; ++i) {
                b.grow(tbox[order[i]]);
                cb.grow(&cent[3 * (size_t)order[i]]);
            }
            nodes[j.node].box = b;
            nodes[j.node].first = j.first;
            nodes[j.node].total = j.count;
            if (j.count == 1) { nodes[j.node].left = j.first; nodes[j.node].count = j.count; continue; }
            // binned SAH over the three axes
            float best_cost = INFINITY;
            int best_axis = -1, best_bin = -1;
            for (int ax = 0; ax < 3; ++ax) {
                float c0 = cb.lo[ax], c1 = cb.hi[ax];
                if (!(c1 > c0)) continue;
                Box bb[kBins];
                int32_t bc[kBins];
                for (int k = 0; k < kBins; ++k) { bb[k].reset(); bc[k] = 0; }
                float scale = kBins / (c1 - c0);
                for (int32_t i = j.first; i < j.first + j.count; ++i) {
                    int32_t t = order[i];
                    int k = std::min(kBins - 1, std::max(0, (int)((cent[3 * (size_t)t + ax] - c0) * scale)));
                    bb[k].grow(tbox[t]);
                    bc[k] += weight.empty() ? 1 : weight[(size_t)t];
                }
                float ra[kBins];
                int32_t rc[kBins];
                Box acc;
                acc.reset();
                int32_t n = 0;
                for (int k = kBins - 1; k > 0; --k) { acc.grow(bb[k]); n += bc[k]; ra[k] = acc.area(); rc[k] = n; }
                acc.reset();
                n = 0;
                for (int k = 0; k < kBins - 1; ++k) {
                    acc.grow(bb[k]);
                    n += bc[k];
                    if (n == 0 || rc[k + 1] == 0) continue;
                    float cost = acc.area() * n + ra[k + 1] * rc[k + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = k; }
                }
            }
            float leaf_cost = b.area() * j.count;
            bool make_leaf = j.count <= max_leaf && (best_axis < 0 || best_cost + b.area() >= leaf_cost);
            if (make_leaf) { nodes[j.node].left = j.first; nodes[j.node].count = j.count; continue; }
            int32_t mid;
            if (best_axis >= 0) {
                float c0 = cb.lo[best_axis], scale = kBins / (cb.hi[best_axis] - c0);
                int32_t *beg = order.data() + j.first, *end = beg + j.count;
                int32_t *m = std::partition(beg, end, [&](int32_t t) {
                    int k = std::min(kBins - 1, std::max(0, (int)((cent[3 * (size_t)t + best_axis] - c0) * scale)));
                    return k <= best_bin;
                });
                mid = (int32_t)(m - order.data());
            } else {
                mid = j.first + j.count / 2;   // coincident centroids
            }
            if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
            int32_t child = (int32_t)nodes.size();
            nodes.push_back(Node2());
            nodes.push_back(Node2());
            nodes[j.node].left = child;
            nodes[j.node].count = 0;
            st.push_back({child, j.first, mid - j.first});
            st.push_back({child + 1, mid, j.first + j.count - mid});
        }
    }
};

inline uint8_t exp_for_extent(double ext) {
    // smallest e with ext / 2^e <= BVH8_QMAX
    if (!(ext > 0.0)) return (uint8_t)1;   // 2^-126
    int e = (int)std::ceil(std::log2(ext / (double)BVH8_QMAX));
    while (std::ldexp((double)BVH8_QMAX, e) < ext) ++e;
    e = std::max(-126, std::min(127, e));
    return (uint8_t)(e + 127);
}

}  // namespace

int host_bvh_build(const float *verts, int64_t n_verts, const int32_t *faces, int64_t n_faces, HostBvh *out) {
    (void)n_verts;
    std::memset(out, 0, sizeof(*out));
    const int64_t F = n_faces;
    Builder B;
    B.tbox.resize((size_t)std::max<int64_t>(F, 1));
    B.cent.resize((size_t)std::max<int64_t>(F, 1) * 3);
    B.order.resize((size_t)F);
    std::vector<TriRecord> recs((size_t)F);
    Box scene;
    scene.reset();
    for (int64_t f = 0; f < F; ++f) {
        const float *a = verts + 3 * (int64_t)faces[3 * f], *b = verts + 3 * (int64_t)faces[3 * f + 1], *c = verts + 3 * (int64_t)faces[3 * f + 2];
        Box tb;
        tb.reset();
        tb.grow(a); tb.grow(b); tb.grow(c);
        TriRecord &r = recs[(size_t)f];
        for (int k = 0; k < 3; ++k) {
            r.v0[k] = a[k];
            r.e1[k] = b[k] - a[k];
            r.e2[k] = c[k] - a[k];
            // the record's own corners (v0+e1, v0+e2 after rounding) must be inside the box too
            float p1 = r.v0[k] + r.e1[k], p2 = r.v0[k] + r.e2[k];
            tb.lo[k] = std::min(tb.lo[k], std::min(p1, p2));
            tb.hi[k] = std::max(tb.hi[k], std::max(p1, p2));
            B.cent[3 * (size_t)f + k] = 0.5f * (tb.lo[k] + tb.hi[k]);
        }
        r.prim = (int32_t)f;
        r.pad[0] = r.pad[1] = 0;
        {   // zero-area triangle (e1 x e2 == 0 exactly): never hit, same rule as oracle/intersect.c
            const float cx = r.e1[1] * r.e2[2] - r.e1[2] * r.e2[1], cy = r.e1[2] * r.e2[0] - r.e1[0] * r.e2[2], cz = r.e1[0] * r.e2[1] - r.e1[1] * r.e2[0];
            if (cx == 0.f && cy == 0.f && cz == 0.f)
                for (int k = 0; k < 3; ++k) r.e1[k] = r.e2[k] = 0.f;
        }
        B.tbox[(size_t)f] = tb;
        B.order[(size_t)f] = (int32_t)f;
        scene.grow(tb);
    }
    if (F == 0) { for (int k = 0; k < 3; ++k) { scene.lo[k] = 0.f; scene.hi[k] = 0.f; } }
    // conservative padding: absorbs fp32 slop of the slab test and of Moller-Trumbore near edges
    float ext = std::max(scene.hi[0] - scene.lo[0], std::max(scene.hi[1] - scene.lo[1], scene.hi[2] - scene.lo[2]));
    float amax = 0.f;
    for (int k = 0; k < 3; ++k) amax = std::max(amax, std::max(std::fabs(scene.lo[k]), std::fabs(scene.hi[k])));
    const float pad = 1e-5f * std::max(ext, amax) + 1e-30f;
    for (int64_t f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) { B.tbox[(size_t)f].lo[k] -= pad; B.tbox[(size_t)f].hi[k] += pad; }

    B.nodes.reserve((size_t)(2 * F + 2));
    B.nodes.push_back(Node2());
    if (F > 0) B.build_range(0, 0, (int32_t)F);
    else { B.nodes[0].box = scene; B.nodes[0].left = 0; B.nodes[0].count = 0; B.nodes[0].first = 0; B.nodes[0].total = 0; }

    // ---- collapse to 8-wide, breadth-first so that the children of a node are contiguous
    std::vector<Bvh8Node> wide;
    std::vector<TriRecord> tris;
    tris.reserve((size_t)F);
    struct Pending { int32_t n2; };
    std::vector<int32_t> queue;   // wide node i corresponds to BVH2 node queue[i]
    std::vector<int32_t> depth;
    wide.push_back(Bvh8Node());
    queue.push_back(0);
    depth.push_back(1);
    int32_t max_depth = 1;
    double sah = 0.0;
    const float root_area = std::max(B.nodes[0].box.area(), 1e-30f);
    for (size_t wi = 0; wi < wide.size(); ++wi) {
        const int32_t n2 = queue[wi];
        int32_t ch[8];
        int nch = 0;
        if (F == 0) {
            nch = 0;
        } else if (B.nodes[n2].total <= kMaxLeaf) {
            ch[nch++] = n2;   // tiny scene: the root itself is a leaf child
        } else {
            ch[nch++] = B.nodes[n2].left;
            ch[nch++] = B.nodes[n2].left + 1;
            while (nch < 8) {
                int best = -1;
                float ba = -1.f;
                for (int i = 0; i < nch; ++i)
                    if (B.nodes[ch[i]].total > kMaxLeaf) {      // a subtree of <= 3 triangles is ONE leaf child, whatever its shape
                        float a = B.nodes[ch[i]].box.area();
                        if (a > ba) { ba = a; best = i; }
                    }
                if (best < 0) break;
                int32_t c = ch[best];
                ch[best] = B.nodes[c].left;
                ch[nch++] = B.nodes[c].left + 1;
            }
        }
        Box nb;
        nb.reset();
        for (int i = 0; i < nch; ++i) nb.grow(B.nodes[ch[i]].box);
        if (nch == 0) nb = scene;
        sah += nb.area() / root_area;
        // ---- slot assignment: greedy minimum of cost[child][slot] = dot(centroid_child - centroid_node, D_slot)
        int slot_of[8], child_in_slot[8];
        for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
        {
            float cost[8][8];
            float nc[3];
            for (int k = 0; k < 3; ++k) nc[k] = 0.5f * (nb.lo[k] + nb.hi[k]);
            for (int i = 0; i < nch; ++i) {
                const Box &cb = B.nodes[ch[i]].box;
                float cc[3];
                for (int k = 0; k < 3; ++k) cc[k] = 0.5f * (cb.lo[k] + cb.hi[k]) - nc[k];
                for (int s = 0; s < 8; ++s) {
                    float dx = (s & 4) ? -1.f : 1.f, dy = (s & 2) ? -1.f : 1.f, dz = (s & 1) ? -1.f : 1.f;
                    cost[i][s] = cc[0] * dx + cc[1] * dy + cc[2] * dz;
                }
                slot_of[i] = -1;
            }
            for (int it = 0; it < nch; ++it) {
                float bc = INFINITY;
                int bi = -1, bs = -1;
                for (int i = 0; i < nch; ++i) {
                    if (slot_of[i] >= 0) continue;
                    for (int s = 0; s < 8; ++s) {
                        if (child_in_slot[s] >= 0) continue;
                        if (cost[i][s] < bc) { bc = cost[i][s]; bi = i; bs = s; }
                    }
                }
                slot_of[bi] = bs;
                child_in_slot[bs] = bi;
            }
        }
        // ---- encode
        Bvh8Node N;
        std::memset(&N, 0, sizeof(N));
        for (int k = 0; k < 3; ++k) {
            N.p[k] = nb.lo[k];
            N.e[k] = exp_for_extent((double)nb.hi[k] - (double)nb.lo[k]);
        }
        N.child_base = (uint32_t)wide.size();
        N.tri_base = (uint32_t)tris.size();
        uint32_t tri_off = 0;
        for (int s = 0; s < 8; ++s) {
            int i = child_in_slot[s];
            if (i < 0) continue;
            const Node2 &c = B.nodes[ch[i]];
            const uint8_t *dummy = nullptr;
            (void)dummy;
            bvh8_q_t *qlo[3] = {N.qlo_x, N.qlo_y, N.qlo_z}, *qhi[3] = {N.qhi_x, N.qhi_y, N.qhi_z};
            for (int k = 0; k < 3; ++k) {
                double sc = std::ldexp(1.0, (int)N.e[k] - 127);
                double lo = std::floor(((double)c.box.lo[k] - (double)N.p[k]) / sc);
                double hi = std::ceil(((double)c.box.hi[k] - (double)N.p[k]) / sc);
                qlo[k][s] = bvh8_encode_q((int)std::max(0.0, std::min((double)BVH8_QMAX, lo)));
                qhi[k][s] = bvh8_encode_q((int)std::max(0.0, std::min((double)BVH8_QMAX, hi)));
            }
            if (c.total > kMaxLeaf) {
                N.imask |= (uint8_t)(1u << s);
                N.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
                wide.push_back(Bvh8Node());
                queue.push_back(ch[i]);
                depth.push_back(depth[wi] + 1);
                max_depth = std::max(max_depth, depth[wi] + 1);
            } else {
                const uint32_t unary = c.total == 1 ? 1u : (c.total == 2 ? 3u : 7u);
                N.meta[s] = (uint8_t)((unary << 5) | tri_off);
                for (int32_t t = 0; t < c.total; ++t) tris.push_back(recs[(size_t)B.order[(size_t)(c.first + t)]]);
                tri_off += (uint32_t)c.total;
            }
        }
        wide[wi] = N;
    }
    out->n_nodes = (int64_t)wide.size();
    out->n_tris = (int64_t)tris.size();
    out->nodes = (Bvh8Node *)std::malloc(sizeof(Bvh8Node) * wide.size());
    out->tris = (TriRecord *)std::malloc(sizeof(TriRecord) * std::max<size_t>(tris.size(), 1));
    if (!out->nodes || !out->tris) return -1;
    std::memcpy(out->nodes, wide.data(), sizeof(Bvh8Node) * wide.size());
    if (!tris.empty()) std::memcpy(out->tris, tris.data(), sizeof(TriRecord) * tris.size());
    for (int k = 0; k < 3; ++k) { out->lo[k] = scene.lo[k]; out->hi[k] = scene.hi[k]; }
    out->sah_cost = (float)sah;
    out->max_depth = max_depth;
    return 0;
}

void host_bvh_free(HostBvh *b) {
    if (!b) return;
    std::free(b->nodes);
    std::free(b->tris);
    b->nodes = nullptr;
    b->tris = nullptr;
}


// Top of the device builder's tree (bvh_device.cuh): binned SAH over the boxes of C clusters (subtrees of the Morton hierarchy with at
// most IRIS_SAH_TREELET primitives -- a few hundred to a few thousand boxes), split down to single clusters.  Returns the clusters in
// tree order (`order`, C entries) and the internal nodes in pre-order: node k covers order[first[k] .. first[k] + count[k]) and splits
// after its first nleft[k] clusters; left[k] / right[k] = index of the child node, or -1 - (position in order) when the child is one cluster.
int host_sah_top(const float *boxes, const int32_t *sizes, int32_t C, int32_t *order, int32_t *first, int32_t *count, int32_t *nleft, int32_t *left, int32_t *right) {
    if (C < 2) return 0;
    Builder B;
    B.max_leaf = 1;
    if (sizes) B.weight.assign(sizes, sizes + C);              // SAH cost = area x PRIMITIVES on each side, not clusters
    B.tbox.resize((size_t)C);
    B.cent.resize((size_t)C * 3);
    B.order.resize((size_t)C);
    for (int32_t c = 0; c < C; ++c) {
        for (int k = 0; k < 3; ++k) {
            B.tbox[(size_t)c].lo[k] = boxes[6 * (size_t)c + k];
            B.tbox[(size_t)c].hi[k] = boxes[6 * (size_t)c + 3 + k];
            B.cent[3 * (size_t)c + k] = 0.5f * (boxes[6 * (size_t)c + k] + boxes[6 * (size_t)c + 3 + k]);
        }
        B.order[(size_t)c] = c;
    }
    B.nodes.reserve((size_t)(2 * C + 2));
    B.nodes.push_back(Node2());
    B.build_range(0, 0, C);
    for (int32_t c = 0; c < C; ++c) order[c] = B.order[(size_t)c];
    // pre-order numbering of the internal nodes
    std::vector<int32_t> id_of(B.nodes.size(), -1), st;
    int32_t n_inner = 0;
    st.push_back(0);
    while (!st.empty()) {
        const int32_t n2 = st.back();
        st.pop_back();
        if (B.nodes[(size_t)n2].count != 0) continue;          // a single cluster
        id_of[(size_t)n2] = n_inner++;
        st.push_back(B.nodes[(size_t)n2].left + 1);
        st.push_back(B.nodes[(size_t)n2].left);
    }
    for (size_t n2 = 0; n2 < B.nodes.size(); ++n2) {
        const int32_t k = id_of[n2];
        if (k < 0) continue;
        const Node2 &nd = B.nodes[n2], &L = B.nodes[(size_t)nd.left], &R = B.nodes[(size_t)nd.left + 1];
        first[k] = nd.first;
        count[k] = nd.total;
        nleft[k] = L.total;
        left[k] = L.count != 0 ? -1 - L.first : id_of[(size_t)nd.left];
        right[k] = R.count != 0 ? -1 - R.first : id_of[(size_t)nd.left + 1];
    }
    return n_inner;                                            // == C - 1
}
