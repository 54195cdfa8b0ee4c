// bvh8.h -- in-HBM layout of the scene: 8-wide compressed BVH nodes + 48-byte triangle records.
//
// Node, default format (IRIS_NODE_FP16 = 0): 80 bytes, five 16-byte words (5 x ld.global.nc.v4)
//   w0: origin p.xyz (fp32) | e.x e.y e.z imask            child boxes are p + q * 2^e, q in [0,255]
//   w1: child_base (u32) | tri_base (u32) | meta[0..3] | meta[4..7]
//   w2: qlo_x[0..7] qlo_y[0..7]      w3: qlo_z[0..7] qhi_x[0..7]      w4: qhi_y[0..7] qhi_z[0..7]
// Wide format (IRIS_NODE_FP16 = 1, compile-time A/B): 128 bytes, q in [0,2047] stored as the fp16 VALUE (integers up to 2048 are exact):
//   w2..w7: qlo_x[0..7], qlo_y, qlo_z, qhi_z, qhi_y, qhi_x
// It removes the byte extraction (24 PRMT) and the near/far SELs from the ALU pipe and quantises the boxes 8x finer, but measured on
// B200 it only wins on coherent rays (camera rays +8 %) and loses on incoherent ones (random rays -15 %, path_tracing_single's
// secondary rays -5 %): with 1.6x the bytes per node fewer nodes stay in L1.  The compact format is therefore the default.
// meta[s]: 0 = empty slot; inner child: 0b001xxxxx with xxxxx = 24 + s; leaf child: unary triangle count in the
// top three bits (001,011,111) and the offset from tri_base in the low five.  Inner children of a node are
// contiguous from child_base in slot order (imask bit s = slot s is inner); triangles of a node are contiguous
// from tri_base.  Children are assigned to slots so that (slot ^ octant) orders them front to back.
//
// Triangle record (three 16-byte words): v0.xyz e1.x | e1.yz e2.xy | e2.z prim pad pad, with e1 = v1 - v0,
// e2 = v2 - v0 rounded once in fp32 (same as oracle/intersect.c) and prim = index into the caller's face array.
#pragma once
#include <stdint.h>

// largest subtree (in triangles) that becomes ONE leaf child slot; the slot's meta byte holds the count in unary (1..3)
#ifndef IRIS_MAX_LEAF
#define IRIS_MAX_LEAF 3
#endif
#ifndef IRIS_NODE_FP16
#define IRIS_NODE_FP16 0
#endif
#if IRIS_NODE_FP16
typedef uint16_t bvh8_q_t;
#define BVH8_QMAX 2047
#define BVH8_NODE_F4 8          // 16-byte words per node
#else
typedef uint8_t bvh8_q_t;
#define BVH8_QMAX 255
#define BVH8_NODE_F4 5
#endif

struct Bvh8Node {
    float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base;
    uint32_t tri_base;
    uint8_t meta[8];
#if IRIS_NODE_FP16
    bvh8_q_t qlo_x[8], qlo_y[8], qlo_z[8], qhi_z[8], qhi_y[8], qhi_x[8];      // words 2..7: near word + far word = 9 on every axis
#else
    bvh8_q_t qlo_x[8], qlo_y[8], qlo_z[8], qhi_x[8], qhi_y[8], qhi_z[8];
#endif
};
static_assert(sizeof(Bvh8Node) == 16 * BVH8_NODE_F4, "Bvh8Node size");

// stored form of the integer plane coordinate v in [0, BVH8_QMAX]: the byte itself, or the fp16 bit pattern of the value v
#if defined(__CUDACC__)
__host__ __device__
#endif
inline bvh8_q_t bvh8_encode_q(int v) {
#if IRIS_NODE_FP16
    if (v <= 0) return 0;
    int e = 0;
    while ((v >> (e + 1)) != 0) ++e;                       // floor(log2 v), v < 2048 -> e <= 10
    return (bvh8_q_t)(((e + 15) << 10) | ((v << (10 - e)) & 0x3FF));
#else
    return (bvh8_q_t)v;
#endif
}
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int bvh8_decode_q(bvh8_q_t q) {
#if IRIS_NODE_FP16
    if (q == 0) return 0;
    const int e = (q >> 10) - 15;
    return (0x400 | (q & 0x3FF)) >> (10 - e);
#else
    return (int)q;
#endif
}

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
// SAH tree over C cluster boxes (6 floats each) for the top of the device builder's tree; see bvh_build.cpp.
int host_sah_top(const float *boxes, const int32_t *sizes, int32_t C, int32_t *order, int32_t *first, int32_t *count, int32_t *nleft, int32_t *left, int32_t *right);
