// TEST INFRASTRUCTURE -- not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load the library this
// file builds (oracle/_ref/libikdtree_ref.so).
//
// Thin C-ABI wrapper around the reference's OWN k-d tree
// (/root/reference/include/ikd-Tree/ikd_Tree.{h,cpp}, compiled unmodified from
// where it lies through oracle/shim/, see oracle/Makefile).  It exposes exactly
// the two calls the hot path makes:
//   KD_TREE::Build          (reference src/laserMapping.cpp:1411-1419)
//   KD_TREE::Nearest_Search (reference src/laserMapping.cpp:1543,
//                            include/ikd-Tree/ikd_Tree.cpp:350-380)
// ikd-Tree returns neighbour POINT COPIES, not indices (ikd_Tree.h:177); we
// stash each map point's index in the bits of its `curvature` field (never
// touched by the tree) so callers can recover indices.
#include "ikd_Tree.h"
#include <omp.h>
#include <cstdint>

extern "C" {

// Heap-allocate and never destroy: KD_TREE embeds a 1M-entry MANUAL_Q
// (ikd_Tree.h:17,82) and its destructor joins a thread whose entry point has no
// `return` (ikd_Tree.cpp:182-185, UB) -- so the tree is deliberately leaked.
void* ikdref_build(const float* xyz, int M, int stride_floats) {
    KD_TREE* tree = new KD_TREE(0.5f, 0.6f, 0.2f);
    PointVector pts((size_t)M);
    for (int i = 0; i < M; ++i) {
        PointType p;
        p.x = xyz[(size_t)i * stride_floats + 0];
        p.y = xyz[(size_t)i * stride_floats + 1];
        p.z = xyz[(size_t)i * stride_floats + 2];
        int32_t idx = i;
        std::memcpy(&p.curvature, &idx, 4);
        pts[i] = p;
    }
    tree->Build(pts);
    return tree;
}

int ikdref_size(void* t) { return static_cast<KD_TREE*>(t)->size(); }

// q: nq x 3 floats.  idx: nq x k (filled with -1 where fewer than k found),
// d2: nq x k (INFINITY where missing), in the ascending order Nearest_Search
// returns.  nthreads mirrors the reference's omp_set_num_threads(MP_PROC_NUM)
// (src/laserMapping.cpp:1516-1518).
int ikdref_knn(void* t, const float* q, int nq, int k, int* idx, float* d2, int nthreads) {
    KD_TREE* tree = static_cast<KD_TREE*>(t);
    if (nthreads < 1) nthreads = 1;
    omp_set_num_threads(nthreads);
#pragma omp parallel for
    for (int i = 0; i < nq; ++i) {
        PointType p;
        p.x = q[3 * (size_t)i + 0];
        p.y = q[3 * (size_t)i + 1];
        p.z = q[3 * (size_t)i + 2];
        PointVector near;
        std::vector<float> dist;
        tree->Nearest_Search(p, k, near, dist);
        for (int j = 0; j < k; ++j) {
            if (j < (int)near.size()) {
                int32_t id;
                std::memcpy(&id, &near[j].curvature, 4);
                idx[(size_t)i * k + j] = id;
                d2[(size_t)i * k + j] = dist[j];
            } else {
                idx[(size_t)i * k + j] = -1;
                d2[(size_t)i * k + j] = INFINITY;
            }
        }
    }
    return 0;
}

}  // extern "C"
