// bvh8.h -- in-HBM layout of the scene: 80-byte 8-wide compressed BVH nodes + 48-byte triangle records.
//
// Node (five 16-byte words, loaded as 5 x ld.global.nc.v4):
//   w0: origin p.xyz (fp32) | e.x e.y e.z imask            child boxes are p + q * 2^e, q in [0,255]
//   w1: child_base (u32) | tri_base (u32) | meta[0..3] | meta[4..7]
//   w2: qlo_x[0..7] qlo_y[0..7]      w3: qlo_z[0..7] qhi_x[0..7]      w4: qhi_y[0..7] qhi_z[0..7]
// meta[s]: 0 = empty slot; inner child: 0b001xxxxx with xxxxx = 24 + s; leaf child: unary triangle count in the
// top three bits (001,011,111) and the offset from tri_base in the low five.  Inner children of a node are
// contiguous from child_base in slot order (imask bit s = slot s is inner); triangles of a node are contiguous
// from tri_base.  Children are assigned to slots so that (slot ^ octant) orders them front to back.
//
// Triangle record (three 16-byte words): v0.xyz e1.x | e1.yz e2.xy | e2.z prim pad pad, with e1 = v1 - v0,
// e2 = v2 - v0 rounded once in fp32 (same as oracle/intersect.c) and prim = index into the caller's face array.
#pragma once
#include <stdint.h>

struct Bvh8Node {
    float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base;
    uint32_t tri_base;
    uint8_t meta[8];
    uint8_t qlo_x[8], qlo_y[8], qlo_z[8], qhi_x[8], qhi_y[8], qhi_z[8];
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

struct TriRecord {
    float v0[3];
    float e1[3];
    float e2[3];
    int32_t prim;
    uint32_t pad[2];
};
static_assert(sizeof(TriRecord) == 48, "TriRecord must be 48 bytes");

struct HostBvh {
    Bvh8Node *nodes = nullptr;
    int64_t n_nodes = 0;
    TriRecord *tris = nullptr;
    int64_t n_tris = 0;
    float lo[3], hi[3];
    float sah_cost = 0.f;
    int32_t max_depth = 0;   // levels of 8-wide nodes; the traversal stack needs one entry per level
};

// Host binned-SAH builder (bvh_build.cpp).  Allocates with malloc; free with host_bvh_free.
int host_bvh_build(const float *verts, int64_t n_verts, const int32_t *faces, int64_t n_faces, HostBvh *out);
void host_bvh_free(HostBvh *b);
