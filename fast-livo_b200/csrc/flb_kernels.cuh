// flb_kernels.cuh -- __global__ kernels of the FAST-LIVO hot path for sm_100a.
//
// Three kernel families:
//   * k_lio_pass      : per-point world transform + exact 5-NN over the device map grid
//                       + float32 QR plane fit (rematch passes) + point-to-plane residual
//                       + Jacobian row + block-level H^T H / H^T z reduction
//   * k_vio_pass      : warp-per-patch photometric residual + d r / d xi over 8x8 pixels
//                       + block-level H^T H / H^T z reduction
//   * k_*_finalize    : fixed-order cross-block reduction, the 18-DoF iterated-ESKF solve,
//                       state [+], convergence / rematch / stop control -- all on device
// plus the map-grid build kernels.  No tensor cores: nothing here is a dense contraction.
//
// Determinism: no floating-point atomics anywhere; every reduction has a fixed order, so
// results are bit-reproducible run to run and identical across ranks.
#pragma once

#include <cuda_runtime.h>
#include "flb_device.cuh"

namespace flb {

// ---------------------------------------------------------------------------------------
// device-resident control blocks
// ---------------------------------------------------------------------------------------
struct LioCtrl {
    // loop state, src/laserMapping.cpp:1472-1473, 1506, 1634-1635, 1700-1731
    int iterCount;
    int rematch_num;
    int nearest_search_en;
    int stop;
    // report
    int passes, knn_passes, n_eff_last, converged_last, status;
    double res_mean_last;
    long long rows_total;
};

struct VioCtrl {
    int level;          // 2,1,0 ; -1 when finished
    int iteration;
    int stop;
    float last_error;   // per level
    float now_error;
    int any_solved;     // G valid
    int passes[3];
    float level_error[3];
    long long rows_total;
    int skipped_last, cov_updated, status;
};

struct LioParamsDev {
    double R_LI[9], t_LI[3];
    double sigma;
    int max_iteration;
    double conv_rot_deg, conv_pos_cm;
};

struct VioParamsDev {
    double Rci[9], Pci[3];
    double sigma;
    int max_iteration;
    float conv_rot_deg, conv_pos_cm;
    int force_all_passes;
};

// Packed sizes: W(W+1)/2 upper-triangular products + W (H^T z) + 2 scalars.
__host__ __device__ constexpr int lio_packed(int W) { return W * (W + 1) / 2 + W + 2; }
constexpr int kVioPacked = 21 + 6 + 2;  // + n_meas, skipped

// ---------------------------------------------------------------------------------------
// warp / block reduction helpers (double, fixed order)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------
// map grid build
// ---------------------------------------------------------------------------------------
__global__ void k_map_cell_ids(const float* __restrict__ xyz, int M, int stride, GridDesc g, unsigned* __restrict__ keys,
                               int* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    int cx = (int)floorf((x - g.ox) * g.inv_cell), cy = (int)floorf((y - g.oy) * g.inv_cell),
        cz = (int)floorf((z - g.oz) * g.inv_cell);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    cz = min(max(cz, 0), g.nz - 1);
    keys[i] = (unsigned)((cz * g.ny + cy) * g.nx + cx);
    vals[i] = i;
}

// After the stable sort by cell: gather points (w = original index bits) and mark cell starts.
__global__ void k_map_gather(const float* __restrict__ xyz, int M, int stride, const unsigned* __restrict__ keys_sorted,
                             const int* __restrict__ vals_sorted, float4* __restrict__ pts, int* __restrict__ cell_start,
                             int ncell) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int src = vals_sorted[i];
    float4 p;
    p.x = xyz[(size_t)src * stride];
    p.y = xyz[(size_t)src * stride + 1];
    p.z = xyz[(size_t)src * stride + 2];
    p.w = __int_as_float(src);
    pts[i] = p;
    const unsigned k = keys_sorted[i];
    const unsigned kprev = (i == 0) ? 0u : keys_sorted[i - 1];
    if (i == 0) {
        for (unsigned c = 0; c <= k; ++c) cell_start[c] = 0;
    } else if (k != kprev) {
        for (unsigned c = kprev + 1; c <= k; ++c) cell_start[c] = i;
    }
    if (i == M - 1) {
        for (unsigned c = k + 1; c <= (unsigned)ncell; ++c) cell_start[c] = M;
    }
}

// ---------------------------------------------------------------------------------------
// Group-cooperative exact 5-NN: kGroup consecutive lanes share one query.  The (z,y) cell rows of
// each Chebyshev ring are dealt round-robin to the lanes (a row's x-range is one contiguous run of
// the sorted map), every lane keeps a private sorted top-5, and the group merges them with
// shuffles after each ring.  Same termination rule and same strict-< semantics as knn5_grid
// (flb_device.cuh); 8x more threads in flight and an ~8x shorter dependent-load chain per query.
// Must be called by all kGroup lanes of the group with identical (qx,qy,qz); every lane returns
// the merged result (sorted map positions).
// ---------------------------------------------------------------------------------------
constexpr int kGroup = 8;

__device__ __forceinline__ void knn5_group_merge(Top5& loc, Top5& G, int sub, unsigned gmask) {
#pragma unroll
    for (int j = 0; j < kMatch; ++j) {
        float md = loc.d[0];
        int mi = loc.i[0];
        int ml = sub;
#pragma unroll
        for (int off = kGroup / 2; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(gmask, md, off);
            const int oi = __shfl_xor_sync(gmask, mi, off);
            const int ol = __shfl_xor_sync(gmask, ml, off);
            if (od < md || (od == md && ol < ml)) { md = od; mi = oi; ml = ol; }
        }
        G.d[j] = md;
        G.i[j] = mi;
        if (ml == sub && mi >= 0) {   // this lane's head won: pop it
#pragma unroll
            for (int q = 0; q < kMatch - 1; ++q) { loc.d[q] = loc.d[q + 1]; loc.i[q] = loc.i[q + 1]; }
            loc.d[kMatch - 1] = INFINITY;
            loc.i[kMatch - 1] = -1;
        }
    }
}

__device__ __forceinline__ void knn_scan_run_thr(const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz,
                                                 float thr, Top5& t) {
    for (int m = s; m < e; ++m) {
        const float4 P = __ldg(pts + m);
        const float d = dist2f(qx, qy, qz, P.x, P.y, P.z);
        if (d < thr) top5_insert(t, d, m);
    }
}

__device__ __forceinline__ void knn5_grid_group(const GridDesc& g, const int* __restrict__ cell_start,
                                                const float4* __restrict__ pts, float qx, float qy, float qz, Top5& G,
                                                int sub, unsigned gmask) {
    top5_init(G);
    const float fx = (qx - g.ox) * g.inv_cell, fy = (qy - g.oy) * g.inv_cell, fz = (qz - g.oz) * g.inv_cell;
    const float lim = (float)(g.max_ring + 2);
    if (!(fx > -lim && fy > -lim && fz > -lim && fx < (float)g.nx + lim && fy < (float)g.ny + lim &&
          fz < (float)g.nz + lim))
        return;
    const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
    const float mx = fminf(fx - (float)cx, (float)(cx + 1) - fx);
    const float my = fminf(fy - (float)cy, (float)(cy + 1) - fy);
    const float mz = fminf(fz - (float)cz, (float)(cz + 1) - fz);
    const float margin = fmaxf(fminf(mx, fminf(my, mz)) * g.cell - 1e-3f * g.cell, 0.f);
    Top5 loc;
    top5_init(loc);
    float thr = INFINITY;   // current global 5th-best distance
    // rings 0 and 1 are scanned together (a 3x3 block of 3-cell x-rows); then one shell per round
    for (int r = 1; r <= g.max_ring; ++r) {
        const int side = 2 * r + 1;
        const int ntask = side * side;
        const int x0 = cx - r, x1 = cx + r;
        const int xa = x0 < 0 ? 0 : x0, xb = x1 >= g.nx ? g.nx - 1 : x1;
        for (int t = sub; t < ntask; t += kGroup) {
            const int dz = t / side - r, dy = t - (t / side) * side - r;
            const int z = cz + dz, y = cy + dy;
            if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) continue;
            const int rowbase = (z * g.ny + y) * g.nx;
            const bool face = (r == 1) || dz == -r || dz == r || dy == -r || dy == r;
            if (face) {
                if (xa <= xb) {
                    const int s = __ldg(cell_start + rowbase + xa), e = __ldg(cell_start + rowbase + xb + 1);
                    knn_scan_run_thr(pts, s, e, qx, qy, qz, thr, loc);
                }
            } else {
                if (x0 >= 0 && x0 < g.nx) {
                    const int s = __ldg(cell_start + rowbase + x0), e = __ldg(cell_start + rowbase + x0 + 1);
                    knn_scan_run_thr(pts, s, e, qx, qy, qz, thr, loc);
                }
                if (x1 >= 0 && x1 < g.nx) {
                    const int s = __ldg(cell_start + rowbase + x1), e = __ldg(cell_start + rowbase + x1 + 1);
                    knn_scan_run_thr(pts, s, e, qx, qy, qz, thr, loc);
                }
            }
        }
        knn5_group_merge(loc, G, sub, gmask);
        // lane 0 carries the merged list forward, the others restart empty below the global threshold
        thr = G.d[kMatch - 1];
        if (sub == 0) loc = G; else top5_init(loc);
        const float bound = (float)r * g.cell + margin;
        const float b2 = bound * bound;
        if (G.d[kMatch - 1] <= b2) break;
        if (b2 > g.max_d2 * 1.0001f) break;
    }
}

// ---------------------------------------------------------------------------------------
// standalone kNN (flb_knn): kGroup lanes per query
// ---------------------------------------------------------------------------------------
__global__ void k_knn(GridDesc g, const int* __restrict__ cell_start, const float4* __restrict__ pts,
                      const float* __restrict__ q, int nq, int* __restrict__ idx, float* __restrict__ d2) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gid / kGroup, sub = threadIdx.x & (kGroup - 1);
    if (i >= nq) return;   // group-uniform
    const unsigned gmask = ((1u << kGroup) - 1u) << ((threadIdx.x & 31) & ~(kGroup - 1));
    Top5 t;
    knn5_grid_group(g, cell_start, pts, q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2], t, sub, gmask);
    if (sub != 0) return;
#pragma unroll
    for (int j = 0; j < kMatch; ++j) {
        const bool ok = (t.i[j] >= 0) && !(t.d[j] > g.max_d2);
        idx[(size_t)i * kMatch + j] = ok ? __float_as_int(__ldg(pts + t.i[j]).w) : -1;
        d2[(size_t)i * kMatch + j] = ok ? t.d[j] : INFINITY;
    }
}

// ---------------------------------------------------------------------------------------
// LIO pass
// ---------------------------------------------------------------------------------------
struct LioArgs {
    // inputs
    const float4* scan;          // N body points
    int N;
    GridDesc grid;
    const int* cell_start;
    const float4* map_pts;
    const State18* state;        // device state (pose read each pass)
    LioParamsDev prm;
    float plane_thr;
    // control: force_rematch < 0 -> read ctrl->nearest_search_en / ctrl->stop
    LioCtrl* ctrl;
    int force_rematch;
    const double* pose_override; // optional 12 doubles (R,p) on device; used by flb_lio_pass
    // per-point persistent
    unsigned char* sel;          // point_selected_surf
    float4* plane;               // cached pabcd
    unsigned char* plane_ok;
    // outputs
    double* partials;            // gridDim.x * lio_packed(W)
    // export (may be null)
    float* x_world;              // N*3
    int* x_nn_idx;               // N*5
    float* x_nn_d2;              // N*5
    float* x_pd2;                // N
    unsigned char* x_rowmask;    // N
    double* x_rows;              // N*W
    double* x_meas;              // N
};

// One scan point.  Rematch passes: called by all kGroup lanes of a group (cooperative kNN), lane
// sub == 0 carries on with the plane fit / residual / row; plain passes: one thread per point (sub = 0).
template <int W>
__device__ __forceinline__ void lio_point(const LioArgs& a, const LioPose& pose, bool rematch, int i, int sub, unsigned gmask,
                                          bool& active, double (&row)[W], double& z, double& absres) {
    active = false;
    z = 0.0;
    absres = 0.0;
    const float4 b4 = __ldg(a.scan + i);
    const float pb[3] = {b4.x, b4.y, b4.z};
    double pI[3];
    float pw[3];
    lio_body_to_world(pose, pb, pI, pw);
    bool sel;
    bool pok;
    float pabcd[4];
    if (rematch) {
        Top5 t;
        knn5_grid_group(a.grid, a.cell_start, a.map_pts, pw[0], pw[1], pw[2], t, sub, gmask);
        if (sub != 0) return;
        const bool five = t.i[kMatch - 1] >= 0;
        sel = five && !(t.d[kMatch - 1] > a.grid.max_d2);     // src/laserMapping.cpp:1549 (+ :1567 size check)
        pok = false;
        float nb[kMatch][3];
        int oidx[kMatch];
#pragma unroll
        for (int j = 0; j < kMatch; ++j) {
            oidx[j] = -1;
            if (t.i[j] >= 0) {
                const float4 P = __ldg(a.map_pts + t.i[j]);
                nb[j][0] = P.x; nb[j][1] = P.y; nb[j][2] = P.z;
                oidx[j] = __float_as_int(P.w);
            } else {
                nb[j][0] = nb[j][1] = nb[j][2] = 0.f;
            }
        }
        pabcd[0] = pabcd[1] = pabcd[2] = pabcd[3] = 0.f;
        if (sel) pok = plane_fit5(nb, a.plane_thr, pabcd);     // :1571
        a.plane[i] = make_float4(pabcd[0], pabcd[1], pabcd[2], pabcd[3]);
        a.plane_ok[i] = pok ? 1 : 0;
        if (a.x_nn_idx) {
#pragma unroll
            for (int j = 0; j < kMatch; ++j) {
                const bool ok = (t.i[j] >= 0) && !(t.d[j] > a.grid.max_d2);
                a.x_nn_idx[(size_t)i * kMatch + j] = ok ? oidx[j] : -1;
                a.x_nn_d2[(size_t)i * kMatch + j] = ok ? t.d[j] : INFINITY;
            }
        }
    } else {
        // through L2: in the persistent kernel these were written by another SM in an earlier pass
        sel = __ldcg(a.sel + i) != 0;
        pok = __ldcg(a.plane_ok + i) != 0;
        const float4 pl = __ldcg(a.plane + i);
        pabcd[0] = pl.x; pabcd[1] = pl.y; pabcd[2] = pl.z; pabcd[3] = pl.w;
    }
    float pd2 = 0.f;
    if (sel) {
        sel = false;                                           // :1570
        if (pok) {
            if (lio_residual(pabcd, pw, pb, &pd2)) sel = true; // :1573-1584
        }
    }
    a.sel[i] = sel ? 1 : 0;
    active = sel && ((double)fabsf(pd2) <= 2.0);               // :1593
    if (active) {
        if (W == 6) lio_row6(pose, pI, pabcd, row);
        else lio_row12(pose, pI, pb, pabcd, row);
        z = -(double)pd2;                                      // :1628
        absres = (double)fabsf(pd2);
    }
    if (a.x_world) {
        a.x_world[3 * (size_t)i] = pw[0]; a.x_world[3 * (size_t)i + 1] = pw[1]; a.x_world[3 * (size_t)i + 2] = pw[2];
    }
    if (a.x_pd2) a.x_pd2[i] = pd2;
    if (a.x_rowmask) a.x_rowmask[i] = active ? 1 : 0;
    if (a.x_rows) {
#pragma unroll
        for (int k = 0; k < W; ++k) a.x_rows[(size_t)i * W + k] = active ? row[k] : 0.0;
        a.x_meas[i] = z;
    }
}

__device__ __forceinline__ void lio_make_pose(const LioArgs& a, LioPose& pose) {
    const double* R = a.pose_override ? a.pose_override : a.state->rot;
    const double* p = a.pose_override ? a.pose_override + 9 : a.state->pos;
    for (int i = 0; i < 9; ++i) { pose.R[i] = R[i]; pose.R_LI[i] = a.prm.R_LI[i]; }
    for (int i = 0; i < 3; ++i) { pose.p[i] = p[i]; pose.t_LI[i] = a.prm.t_LI[i]; }
    m3_T(pose.R, pose.Rt);
    m3_T(pose.R_LI, pose.RLIt);
}

// kGroup lanes per scan point (cooperative kNN on rematch passes; lanes sub != 0 idle on plain passes);
// block partial sums written to partials[blockIdx.x][*].
template <int W, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_lio_pass(LioArgs a) {
    constexpr int K = lio_packed(W);
    constexpr int NW = BLOCK / 32;
    __shared__ LioPose s_pose;
    __shared__ double s_acc[NW][K];
    int rematch;
    if (a.force_rematch >= 0) {
        rematch = a.force_rematch;
    } else {
        if (a.ctrl->stop) return;
        rematch = a.ctrl->nearest_search_en;
    }
    if (threadIdx.x == 0) lio_make_pose(a, s_pose);
    __syncthreads();
    const int i = (blockIdx.x * BLOCK + threadIdx.x) / kGroup, sub = threadIdx.x & (kGroup - 1);
    const unsigned gmask = ((1u << kGroup) - 1u) << ((threadIdx.x & 31) & ~(kGroup - 1));
    bool active = false;
    double row[W], z = 0.0, absres = 0.0;
#pragma unroll
    for (int k = 0; k < W; ++k) row[k] = 0.0;
    if (i < a.N && (rematch || sub == 0)) lio_point<W>(a, s_pose, rematch != 0, i, sub, gmask, active, row, z, absres);
    // warp-level reduction of the packed products, then across warps in fixed order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int k = 0;
#pragma unroll
    for (int r = 0; r < W; ++r) {
#pragma unroll
        for (int c = r; c < W; ++c) {
            const double v = warp_sum(active ? row[r] * row[c] : 0.0);
            if (lane == 0) s_acc[warp][k] = v;
            ++k;
        }
    }
#pragma unroll
    for (int r = 0; r < W; ++r) {
        const double v = warp_sum(active ? row[r] * z : 0.0);
        if (lane == 0) s_acc[warp][k] = v;
        ++k;
    }
    {
        const double v = warp_sum(active ? 1.0 : 0.0);
        if (lane == 0) s_acc[warp][k] = v;
        ++k;
        const double v2 = warp_sum(absres);
        if (lane == 0) s_acc[warp][k] = v2;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < K; q += BLOCK) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += s_acc[w][q];
        a.partials[(size_t)blockIdx.x * K + q] = s;
    }
}

// ---------------------------------------------------------------------------------------
// In-warp Gauss-Jordan on an [n x (n+m)] augmented system held in shared memory
// (row stride LD).  lane = row; no pivoting: the systems here are symmetric positive
// definite (H^T H + prior information, or a covariance), for which elimination without
// pivoting is backward stable.  On exit the right m columns hold A^-1 * RHS.
// Returns false (to every lane) on a non-positive / non-finite pivot.
// ---------------------------------------------------------------------------------------
template <int N, int M, int LD>
__device__ __forceinline__ bool warp_gauss_jordan(double* S, int lane) {
    bool ok = true;
    for (int k = 0; k < N; ++k) {
        const double piv = S[k * LD + k];
        if (!(fabs(piv) > 1e-300) || !isfinite(piv)) ok = false;
        __syncwarp();
        if (lane == k) {
            const double inv = 1.0 / piv;
            for (int j = 0; j < N + M; ++j) S[k * LD + j] = S[k * LD + j] * inv;
        }
        __syncwarp();
        if (lane < N && lane != k) {
            const double f = S[lane * LD + k];
            for (int j = 0; j < N + M; ++j) S[lane * LD + j] = S[lane * LD + j] - f * S[k * LD + j];
        }
        __syncwarp();
    }
    return ok;
}

// (cov / sigma)^-1, once per update (the covariance only changes on the stopping pass,
// src/laserMapping.cpp:1715 / src/lidar_selection.cpp:980).  One warp.
__device__ __forceinline__ bool prior_information(const double* cov, double sigma, double* Pinv /*18x18 global*/,
                                                  double* S /*smem 18x36*/, int lane) {
    constexpr int LD = 2 * kDim;
    if (lane < kDim) {
        for (int j = 0; j < kDim; ++j) {
            S[lane * LD + j] = cov[lane * kDim + j] / sigma;
            S[lane * LD + kDim + j] = (j == lane) ? 1.0 : 0.0;
        }
    }
    __syncwarp();
    const bool ok = warp_gauss_jordan<kDim, kDim, LD>(S, lane);
    if (lane < kDim)
        for (int j = 0; j < kDim; ++j) Pinv[lane * kDim + j] = S[lane * LD + kDim + j];
    __syncwarp();
    return ok;
}

// The shared 18-DoF step (src/laserMapping.cpp:1664-1672, src/lidar_selection.cpp:871-878):
//   K1 = (H_T_H + Pinv)^-1 ; Gc = K1[:, :6]*HTH6 ; sol = sign*K1[:, :6]*HTz + vec - Gc*vec[:6]
// One warp; HTH6/HTz in smem; returns sol[18] and Gc[18][6] in smem.
struct SolveSmem {
    double S[kDim * (kDim + 6)];
    double HTH[36];
    double HTz[6];
    double vec[kDim];
    double sol[kDim];
    double Gc[kDim * 6];
};

__device__ __forceinline__ bool ekf18_solve(SolveSmem& sm, const double* Pinv, const State18* x, const State18* x_prop,
                                            double sign, int lane) {
    constexpr int LD = kDim + 6;
    if (lane < kDim) {
        for (int j = 0; j < kDim; ++j) {
            double v = Pinv[lane * kDim + j];
            if (lane < 6 && j < 6) v = sm.HTH[lane * 6 + j] + v;
            sm.S[lane * LD + j] = v;
        }
        for (int j = 0; j < 6; ++j) sm.S[lane * LD + kDim + j] = (j == lane) ? 1.0 : 0.0;
    }
    if (lane == 31) state_boxminus(*x_prop, *x, sm.vec);   // vec = state_propagat - state
    __syncwarp();
    const bool ok = warp_gauss_jordan<kDim, 6, LD>(sm.S, lane);
    if (lane < kDim) {
        double g[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s += sm.S[lane * LD + kDim + k] * sm.HTH[k * 6 + j];
            g[j] = s;
            sm.Gc[lane * 6 + j] = s;
        }
        double kz = 0.0, gv = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            kz += sm.S[lane * LD + kDim + k] * sm.HTz[k];
            gv += g[k] * sm.vec[k];
        }
        sm.sol[lane] = sign * kz + sm.vec[lane] - gv;
    }
    __syncwarp();
    return ok;
}

// cov <- cov - Gc * cov[:6, :]   ( == (I - G) * cov with G's only non-zero columns 0..5 )
__device__ __forceinline__ void cov_update(double* cov, const double* Gc, double* top /*smem 6x18*/, int lane) {
    for (int e = lane; e < 6 * kDim; e += 32) top[e] = cov[e];
    __syncwarp();
    for (int e = lane; e < kDim * kDim; e += 32) {
        const int i = e / kDim, j = e % kDim;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += Gc[i * 6 + k] * top[k * kDim + j];
        cov[e] = cov[e] - s;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// LIO begin / finalize (single warp each)
// ---------------------------------------------------------------------------------------
struct LioSolveArgs {
    State18* state;
    const State18* state_prop;
    LioCtrl* ctrl;
    double* Pinv;            // 18x18
    const double* partials;  // nblocks * lio_packed(6)   (or 1 x packed after an all-reduce)
    int nblocks;
    LioParamsDev prm;
    double* packed_out;      // optional: lio_packed(6) reduced sums (for flb_lio_pass / all-reduce)
};

__global__ void __launch_bounds__(32) k_lio_begin(LioSolveArgs a) {
    __shared__ double S[kDim * 2 * kDim];
    const int lane = threadIdx.x;
    const bool ok = prior_information(a.state->cov, a.prm.sigma, a.Pinv, S, lane);
    if (lane == 0) {
        LioCtrl c;
        c.iterCount = -1;
        c.rematch_num = 0;
        c.nearest_search_en = 1;
        c.stop = 0;
        c.passes = c.knn_passes = c.n_eff_last = c.converged_last = 0;
        c.status = ok ? 0 : -5;
        c.res_mean_last = 0.0;
        c.rows_total = 0;
        if (!ok) c.stop = 1;
        *a.ctrl = c;
    }
}

// Sum block partials in block order (fixed), lanes across the packed entries.
template <int K>
__device__ __forceinline__ void reduce_partials(const double* partials, int nblocks, double* out /*smem K*/, int lane) {
    for (int q = lane; q < K; q += 32) {
        double s = 0.0;
        int b = 0;
        for (; b + 8 <= nblocks; b += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = partials[(size_t)(b + u) * K + q];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; b < nblocks; ++b) s += partials[(size_t)b * K + q];
        out[q] = s;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32) k_reduce_only(const double* partials, int nblocks, int K, double* out) {
    for (int q = threadIdx.x; q < K; q += 32) {
        double s = 0.0;
        int b = 0;
        for (; b + 8 <= nblocks; b += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = partials[(size_t)(b + u) * K + q];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; b < nblocks; ++b) s += partials[(size_t)b * K + q];
        out[q] = s;
    }
}

__device__ __forceinline__ void unpack_sym6(const double* packed, double* HTH, double* HTz) {
    int k = 0;
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) {
            HTH[r * 6 + c] = packed[k];
            HTH[c * 6 + r] = packed[k];
            ++k;
        }
    for (int r = 0; r < 6; ++r) HTz[r] = packed[k + r];
}

__global__ void __launch_bounds__(32) k_lio_finalize(LioSolveArgs a) {
    constexpr int K = lio_packed(6);
    __shared__ SolveSmem sm;
    __shared__ double packed[K];
    __shared__ double top[6 * kDim];
    const int lane = threadIdx.x;
    if (a.ctrl->stop) return;
    reduce_partials<K>(a.partials, a.nblocks, packed, lane);
    if (lane == 0) unpack_sym6(packed, sm.HTH, sm.HTz);
    __syncwarp();
    const bool ok = ekf18_solve(sm, a.Pinv, a.state, a.state_prop, +1.0, lane);
    const int T = a.prm.max_iteration;
    // control (lane 0), src/laserMapping.cpp:1683-1731
    __shared__ int s_stop;
    if (lane == 0) {
        LioCtrl c = *a.ctrl;
        const int n_eff = (int)packed[K - 2];
        c.passes += 1;
        c.knn_passes += c.nearest_search_en ? 1 : 0;
        c.n_eff_last = n_eff;
        c.res_mean_last = packed[K - 1] / (double)n_eff;
        c.rows_total += n_eff;
        State18* x = a.state;
        state_boxplus(*x, sm.sol);                                                       // :1683
        bool converged = false;
        if ((norm3(sm.sol) * 57.3 < a.prm.conv_rot_deg) && (norm3(sm.sol + 3) * 100 < a.prm.conv_pos_cm))
            converged = true;                                                            // :1688
        c.nearest_search_en = 0;                                                         // :1700
        if (converged || ((c.rematch_num == 0) && (c.iterCount == (T - 2)))) {           // :1701
            c.nearest_search_en = 1;
            c.rematch_num++;
        }
        int stop = 0;
        if (c.rematch_num >= 2 || (c.iterCount == T - 1)) stop = 1;                      // :1708
        if (!ok) { stop = 1; c.status = -5; }
        c.converged_last = converged ? 1 : 0;
        c.iterCount += 1;
        c.stop = stop;
        *a.ctrl = c;
        s_stop = stop && ok;
    }
    __syncwarp();
    if (s_stop) cov_update(a.state->cov, sm.Gc, top, lane);                               // :1715
}

// ---------------------------------------------------------------------------------------
// VIO pass: warp per patch, 2 pixels per lane
// ---------------------------------------------------------------------------------------
struct VioArgs {
    const unsigned char* img;    // continuous, stride == width
    CamModel cam;
    const double* pos;           // Pn*3
    const float* patch;          // Pn*192
    const int* search_level;     // Pn
    int Pn;
    const State18* state;
    const double* pose_override; // 12 doubles or null
    VioParamsDev prm;
    VioCtrl* ctrl;
    int force_level;             // >= 0 : use this level and ignore ctrl
    float* errors;               // Pn  (sub_sparse_map->errors)
    double* partials;            // gridDim.x * kVioPacked
    double* x_z;                 // Pn*64 or null
    double* x_H;                 // Pn*64*6 or null
};

struct LatView {
    const float* p;
    __device__ __forceinline__ float operator()(int r, int c) const { return p[r * 11 + c]; }
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_vio_pass(VioArgs a) {
    constexpr int NW = BLOCK / 32;
    __shared__ VioPose s_pose;
    __shared__ float s_lat[NW][128];
    __shared__ double s_res[NW][64];
    __shared__ double s_acc[NW][kVioPacked];
    int level;
    if (a.force_level >= 0) {
        level = a.force_level;
    } else {
        if (a.ctrl->stop) return;
        level = a.ctrl->level;
    }
    if (threadIdx.x == 0) {
        const double* R = a.pose_override ? a.pose_override : a.state->rot;
        const double* p = a.pose_override ? a.pose_override + 9 : a.state->pos;
        vio_make_pose(a.prm.Rci, a.prm.Pci, R, p, s_pose);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * NW + warp;
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.0;
    double n_meas = 0.0, skipped = 0.0;
    if (i < a.Pn) {
        PatchGeom g;
        const double pos[3] = {a.pos[3 * (size_t)i], a.pos[3 * (size_t)i + 1], a.pos[3 * (size_t)i + 2]};
        vio_patch_geom(a.cam, s_pose, pos, level, a.search_level[i], g);
        if (g.valid) {
            // stage the 11x11 tap lattice (stride = scale px) as float
            const int W = a.cam.width;
            const unsigned char* base = a.img + (size_t)(g.v_i - 5 * g.scale) * W + (g.u_i - 5 * g.scale);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = lane + 32 * q;
                if (e < 121) {
                    const int r = e / 11, c = e - r * 11;
                    s_lat[warp][e] = (float)__ldg(base + (size_t)r * g.scale * W + c * g.scale);
                }
            }
            __syncwarp();
            LatView L{s_lat[warp]};
            const int x = lane >> 2, y0 = (lane & 3) * 2;
            const float* P = a.patch + (size_t)i * 192 + 64 * level;   // P[patch_size_total*level + x*8 + y]
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int y = y0 + q;
                double row[6], res;
                vio_pixel(L, g, s_pose, x, y, __ldg(P + x * 8 + y), row, &res);
                s_res[warp][x * 8 + y] = res;
                int k = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int c = r; c < 6; ++c) { acc[k] += row[r] * row[c]; ++k; }
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) acc[21 + r] += row[r] * res;
                if (a.x_z) {
                    a.x_z[(size_t)i * 64 + x * 8 + y] = res;
#pragma unroll
                    for (int r = 0; r < 6; ++r) a.x_H[((size_t)i * 64 + x * 8 + y) * 6 + r] = row[r];
                }
            }
            __syncwarp();
            if (lane == 0) {
                // patch_error += res*res : float accumulator, double addend (:843) -- sequential, exact
                float pe = 0.0f;
                for (int e = 0; e < 64; ++e) pe = (float)((double)pe + s_res[warp][e] * s_res[warp][e]);
                a.errors[i] = pe;                                   // :851
                n_meas = 64.0;
            }
        } else {
            if (lane == 0) { a.errors[i] = 0.0f; skipped = 1.0; }
            if (a.x_z) {
                for (int e = lane; e < 64; e += 32) {
                    a.x_z[(size_t)i * 64 + e] = 0.0;
                    for (int r = 0; r < 6; ++r) a.x_H[((size_t)i * 64 + e) * 6 + r] = 0.0;
                }
            }
        }
    }
    // warp reduce, then across the block's warps in fixed order
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) s_acc[warp][k] = v;
    }
    if (lane == 0) { s_acc[warp][27] = n_meas; s_acc[warp][28] = skipped; }
    __syncthreads();
    for (int q = threadIdx.x; q < kVioPacked; q += BLOCK) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += s_acc[w][q];
        a.partials[(size_t)blockIdx.x * kVioPacked + q] = s;
    }
}

// ---------------------------------------------------------------------------------------
// VIO begin / finalize
// ---------------------------------------------------------------------------------------
struct VioSolveArgs {
    State18* state;
    const State18* state_prop;
    State18* old_state;
    VioCtrl* ctrl;
    double* Pinv;
    double* G_last;          // 18x6, the G of the last accepted solve (:874, used at :980)
    const double* partials;
    int nblocks;
    const float* errors;     // all patches, patch order
    int Pn_total;
    VioParamsDev prm;
};

__global__ void __launch_bounds__(32) k_vio_begin(VioSolveArgs a) {
    __shared__ double S[kDim * 2 * kDim];
    const int lane = threadIdx.x;
    const bool ok = prior_information(a.state->cov, a.prm.sigma, a.Pinv, S, lane);
    // old_state = *state (:747)
    {
        const double* s = reinterpret_cast<const double*>(a.state);
        double* d = reinterpret_cast<double*>(a.old_state);
        for (int e = lane; e < (int)(sizeof(State18) / sizeof(double)); e += 32) d[e] = s[e];
    }
    if (lane == 0) {
        VioCtrl c;
        c.level = 2;
        c.iteration = 0;
        c.stop = (a.Pn_total == 0 || !ok) ? 1 : 0;       // :969-970
        c.last_error = 1e10f;                            // :971
        c.now_error = 1e10f;
        c.any_solved = 0;
        for (int l = 0; l < 3; ++l) { c.passes[l] = 0; c.level_error[l] = 1e10f; }
        c.rows_total = 0;
        c.skipped_last = 0;
        c.cov_updated = 0;
        c.status = ok ? 0 : -5;
        *a.ctrl = c;
    }
}

// 64 threads: warp 0 reduces + solves (speculatively), warp 1 lane 0 forms the exact
// sequential float sum of the per-patch errors (:852) in patch order.
__global__ void __launch_bounds__(64) k_vio_finalize(VioSolveArgs a) {
    __shared__ SolveSmem sm;
    __shared__ double packed[kVioPacked];
    __shared__ double top[6 * kDim];
    __shared__ float s_error;
    __shared__ int s_flags[2];   // [0] accept, [1] do cov update
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.ctrl->stop) return;
    bool ok = true;
    if (warp == 0) {
        reduce_partials<kVioPacked>(a.partials, a.nblocks, packed, lane);
        if (lane == 0) unpack_sym6(packed, sm.HTH, sm.HTz);
        __syncwarp();
        ok = ekf18_solve(sm, a.Pinv, a.state, a.state_prop, -1.0, lane);     // :871-878 (sign: :878)
    } else if (lane == 0) {
        float e = 0.0f;
        const float* er = a.errors;
        int i = 0;
        for (; i + 8 <= a.Pn_total; i += 8) {
            const float e0 = er[i], e1 = er[i + 1], e2 = er[i + 2], e3 = er[i + 3], e4 = er[i + 4], e5 = er[i + 5],
                        e6 = er[i + 6], e7 = er[i + 7];
            e = e + e0; e = e + e1; e = e + e2; e = e + e3; e = e + e4; e = e + e5; e = e + e6; e = e + e7;
        }
        for (; i < a.Pn_total; ++i) e = e + er[i];
        s_error = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        VioCtrl c = *a.ctrl;
        const long long n_meas = (long long)packed[27];
        const float error = s_error / (float)(unsigned long long)n_meas;                  // :857
        const int level = c.level;
        c.passes[level] += 1;
        c.rows_total += n_meas;
        c.skipped_last = (int)packed[28];
        bool EKF_end = false;
        int accept = 0;
        if (a.prm.force_all_passes || error <= c.last_error) {                            // :861
            accept = 1;
            c.last_error = error;
            if (!a.prm.force_all_passes && (norm3(sm.sol) * 57.3f < a.prm.conv_rot_deg) &&
                (norm3(sm.sol + 3) * 100.0f < a.prm.conv_pos_cm))
                EKF_end = true;                                                           // :883
            c.any_solved = 1;
        } else {
            EKF_end = true;                                                               // :890-891
        }
        if (!ok) { EKF_end = true; c.status = -5; }
        c.iteration += 1;
        int docov = 0;
        if (EKF_end || c.iteration >= a.prm.max_iteration) {
            // level finished -> ComputeJ advances (:974-977)
            c.level_error[level] = c.last_error;
            c.now_error = c.last_error;
            c.level = level - 1;
            c.iteration = 0;
            c.last_error = 1e10f;
            if (c.level < 0 || !ok) {
                c.stop = 1;
                if (c.now_error < 1e10f && ok) { docov = 1; c.cov_updated = 1; }          // :978-981
            }
        }
        *a.ctrl = c;
        s_flags[0] = accept;
        s_flags[1] = docov;
        s_flags[0] |= (c.iteration == 0) ? 2 : 0;   // bit1: a new level starts -> old_state = state
    }
    __syncthreads();
    const int accept = s_flags[0] & 1, newlevel = s_flags[0] & 2, docov = s_flags[1];
    if (warp == 0) {
        double* xs = reinterpret_cast<double*>(a.state);
        double* os = reinterpret_cast<double*>(a.old_state);
        constexpr int NS = (int)(sizeof(State18) / sizeof(double));
        if (accept) {
            for (int e = lane; e < NS; e += 32) os[e] = xs[e];                            // old_state = *state (:863)
            for (int e = lane; e < kDim * 6; e += 32) a.G_last[e] = sm.Gc[e];
            __syncwarp();
            if (lane == 0) state_boxplus(*a.state, sm.sol);                               // :879
        } else {
            for (int e = lane; e < NS; e += 32) xs[e] = os[e];                            // *state = old_state (:890)
        }
        __syncwarp();
        if (newlevel) {
            for (int e = lane; e < NS; e += 32) os[e] = xs[e];                            // :747 of the next level
        }
        __syncwarp();
        if (docov) cov_update(a.state->cov, a.G_last, top, lane);                         // :980
    }
}

}  // namespace flb

// =======================================================================================
// Persistent (one launch per update) kernels
// =======================================================================================
// All blocks are co-resident (cooperative launch, grid <= SMs x occupancy).  Every pass ends
// in a grid barrier whose LAST ARRIVER becomes the leader: it reduces the block partials in
// fixed order, runs the 18-DoF solve with the whole block, publishes state + control through
// L2, and releases the other blocks.  There is no host round trip, no empty launch and no
// cross-block floating-point atomic; the result is bit-identical to the kernel-per-pass path.
namespace flb {

struct GridBarrier {
    unsigned int count;
    unsigned int gen;
    int timeout;
    int pad;
};

// Optional device-side pass trace (profiling aid): %globaltimer (ns) at kernel entry [0], then for
// pass k: [1+2k] = all blocks arrived (leader elected), [2+2k] = leader released the barrier.
constexpr int kTraceLen = 128;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Arrive; returns true (to every thread of the block) when this block is the last arriver.
__device__ __forceinline__ bool grid_arrive(GridBarrier* b, unsigned nblocks, unsigned* s_scratch /*2 words smem*/) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned g = ld_acquire_u32(&b->gen);
        __threadfence();
        const unsigned ticket = atomicAdd(&b->count, 1u);
        __threadfence();
        s_scratch[0] = (ticket == nblocks - 1) ? 1u : 0u;
        s_scratch[1] = g;
    }
    __syncthreads();
    return s_scratch[0] != 0;
}
// Leader: open the barrier.  Call with the whole block.
__device__ __forceinline__ void grid_release(GridBarrier* b, const unsigned* s_scratch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        b->count = 0;
        __threadfence();
        st_release_u32(&b->gen, s_scratch[1] + 1u);
    }
    __syncthreads();
}
// Non-leader: wait.  A watchdog (~seconds) turns a would-be hang into FLB_ERR_TIMEOUT.
__device__ __forceinline__ bool grid_wait(GridBarrier* b, const unsigned* s_scratch) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        unsigned long long spins = 0;
        while (ld_acquire_u32(&b->gen) == s_scratch[1]) {
            __nanosleep(40);
            if (++spins > 3000000ull) { ok = 0; b->timeout = 1; break; }
            if ((spins & 0xffff) == 0 && *((volatile int*)&b->timeout)) { ok = 0; break; }
        }
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Block-parallel Gauss-Jordan on S[N x (N+M)], row stride LD (odd: conflict-free), executed by
// the first `nt` threads of the block (all of them must call; they sync on named barrier 1).
// Same arithmetic as warp_gauss_jordan: scale the pivot row by 1/pivot, then eliminate.
template <int N, int M, int LD, int NT>
__device__ __forceinline__ bool block_gauss_jordan(double* S, int tid) {
    constexpr int COLS = N + M;
    constexpr int E = N * COLS;
    constexpr int PER = (E + NT - 1) / NT;   // elements per thread
    bool ok = true;
    for (int k = 0; k < N; ++k) {
        const double piv = S[k * LD + k];
        if (!(fabs(piv) > 1e-300) || !isfinite(piv)) ok = false;
        const double inv = 1.0 / piv;
        double v[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = tid + q * NT;
            if (e < E) {
                const int i = e / COLS, j = e - i * COLS;
                const double rkj = S[k * LD + j] * inv;
                v[q] = (i == k) ? rkj : S[i * LD + j] - S[i * LD + k] * rkj;
            }
        }
        named_sync(1, NT);
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int e = tid + q * NT;
            if (e < E) {
                const int i = e / COLS, j = e - i * COLS;
                S[i * LD + j] = v[q];
            }
        }
        named_sync(1, NT);
    }
    return ok;
}

struct LeaderSmem {
    State18 x, xp;                       // L2 copies of state / prior
    double S[kDim * 37];                 // GJ workspace (LD = 37 for 18x36, LD = 25 for 18x24)
    double Pinv[kDim * kDim];
    double HTH[36], HTz[6];
    double vec[kDim], sol[kDim], Gc[kDim * 6];
    double top[6 * kDim];
    double packed[32];
    double part[4][32];
    int flags[4];
};

__device__ __forceinline__ void load_state_l2(State18* dst, const State18* src, int tid, int nt) {
    const double* s = reinterpret_cast<const double*>(src);
    double* d = reinterpret_cast<double*>(dst);
    for (int e = tid; e < (int)(sizeof(State18) / sizeof(double)); e += nt) d[e] = __ldcg(s + e);
}
__device__ __forceinline__ void store_state(State18* dst, const State18* src, int tid, int nt, bool with_cov) {
    const double* s = reinterpret_cast<const double*>(src);
    double* d = reinterpret_cast<double*>(dst);
    const int n = with_cov ? (int)(sizeof(State18) / sizeof(double)) : 24;
    for (int e = tid; e < n; e += nt) d[e] = s[e];
}

// Fixed-order reduction of the block partials by the leader: 4 contiguous quarters per entry,
// combined ((q0+q1)+q2)+q3.  (Different association than the kernel-per-pass path's single
// sequential sum: both are deterministic; they agree to ~1e-16 relative.)
template <int K>
__device__ __forceinline__ void leader_reduce(const double* partials, int nblocks, LeaderSmem& sm, int tid) {
    const int q = tid & 31, part = tid >> 5;
    if (part < 4 && q < K) {
        const int per = (nblocks + 3) / 4;
        const int b0 = part * per, b1 = min(nblocks, b0 + per);
        double s = 0.0;
        // batches of 8 independent L2 loads, then the adds in block order: the chain costs one L2
        // round trip per 8 blocks instead of one per block
        int b = b0;
        for (; b + 8 <= b1; b += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(partials + (size_t)(b + u) * K + q);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; b < b1; ++b) s += __ldcg(partials + (size_t)b * K + q);
        sm.part[part][q] = s;
    }
    __syncthreads();
    if (tid < K) sm.packed[tid] = ((sm.part[0][tid] + sm.part[1][tid]) + sm.part[2][tid]) + sm.part[3][tid];
    __syncthreads();
}

// Solve with the whole block (nt_gj threads run the elimination).  Requires sm.x / sm.xp /
// sm.HTH / sm.HTz loaded; `first` computes the prior information from sm.x.cov and stores it
// to global Pinv, otherwise loads it from L2.
template <int NT>
__device__ __forceinline__ bool leader_solve(LeaderSmem& sm, double* Pinv_g, double sigma, double sign, bool first, int tid) {
    constexpr int nt_gj = NT;
    bool ok = true;
    if (tid == nt_gj) state_boxminus(sm.xp, sm.x, sm.vec);      // a thread outside the GJ team
    if (first) {
        constexpr int LD = 37;
        if (tid < nt_gj) {
            for (int e = tid; e < kDim * 2 * kDim; e += nt_gj) {
                const int i = e / (2 * kDim), j = e - i * 2 * kDim;
                sm.S[i * LD + j] = (j < kDim) ? sm.x.cov[i * kDim + j] / sigma : ((j - kDim == i) ? 1.0 : 0.0);
            }
            named_sync(1, nt_gj);
            ok = block_gauss_jordan<kDim, kDim, LD, NT>(sm.S, tid);
            for (int e = tid; e < kDim * kDim; e += nt_gj) {
                const int i = e / kDim, j = e - i * kDim;
                const double v = sm.S[i * LD + kDim + j];
                sm.Pinv[e] = v;
                Pinv_g[e] = v;
            }
        }
    } else {
        for (int e = tid; e < kDim * kDim; e += blockDim.x) sm.Pinv[e] = __ldcg(Pinv_g + e);
    }
    __syncthreads();
    constexpr int LD2 = 25;
    if (tid < nt_gj) {
        for (int e = tid; e < kDim * (kDim + 6); e += nt_gj) {
            const int i = e / (kDim + 6), j = e - i * (kDim + 6);
            double v;
            if (j < kDim) {
                v = sm.Pinv[i * kDim + j];
                if (i < 6 && j < 6) v = sm.HTH[i * 6 + j] + v;
            } else {
                v = (j - kDim == i) ? 1.0 : 0.0;
            }
            sm.S[i * LD2 + j] = v;
        }
        named_sync(1, nt_gj);
        ok = block_gauss_jordan<kDim, 6, LD2, NT>(sm.S, tid) && ok;
    }
    __syncthreads();
    if (tid < kDim * 6) {
        const int i = tid / 6, j = tid - i * 6;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += sm.S[i * LD2 + kDim + k] * sm.HTH[k * 6 + j];
        sm.Gc[tid] = s;
    }
    __syncthreads();
    if (tid < kDim) {
        double kz = 0.0, gv = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            kz += sm.S[tid * LD2 + kDim + k] * sm.HTz[k];
            gv += sm.Gc[tid * 6 + k] * sm.vec[k];
        }
        sm.sol[tid] = sign * kz + sm.vec[tid] - gv;
    }
    __syncthreads();
    return ok;
}

// cov (smem state) <- cov - Gc * cov[:6,:]
__device__ __forceinline__ void leader_cov_update(LeaderSmem& sm, const double* Gc, int tid, int nt) {
    for (int e = tid; e < 6 * kDim; e += nt) sm.top[e] = sm.x.cov[e];
    __syncthreads();
    for (int e = tid; e < kDim * kDim; e += nt) {
        const int i = e / kDim, j = e - i * kDim;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += Gc[i * 6 + k] * sm.top[k * kDim + j];
        sm.x.cov[e] = sm.x.cov[e] - s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// LIO: whole iterated update in one launch
// ---------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_lio_update_persistent(LioArgs a, LioSolveArgs s, GridBarrier* bar,
                                                                 unsigned long long* trace) {
    constexpr int K = lio_packed(6);
    constexpr int NW = BLOCK / 32;
    __shared__ LioPose s_pose;
    __shared__ double s_acc[NW][K];
    __shared__ unsigned s_bar[2];
    __shared__ LeaderSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = s.prm.max_iteration;
    // block-local mirror of the loop state (src/laserMapping.cpp:1472-1473, :1506)
    int iterCount = -1, rematch_num = 0, nearest = 1;
    bool first = true;
    int pass_no = 0;
    if (trace && blockIdx.x == 0 && tid == 0) trace[0] = global_ns();
    for (;;) {
        if (tid == 0) {
            LioPose& pose = s_pose;
            for (int i = 0; i < 9; ++i) { pose.R[i] = __ldcg(&a.state->rot[i]); pose.R_LI[i] = a.prm.R_LI[i]; }
            for (int i = 0; i < 3; ++i) { pose.p[i] = __ldcg(&a.state->pos[i]); pose.t_LI[i] = a.prm.t_LI[i]; }
            m3_T(pose.R, pose.Rt);
            m3_T(pose.R_LI, pose.RLIt);
        }
        __syncthreads();
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        // rematch passes: kGroup lanes per point (cooperative kNN); plain passes: one thread per point
        const int gthreads = gridDim.x * BLOCK, gtid = blockIdx.x * BLOCK + tid;
        const int per = nearest ? kGroup : 1;
        const int sub = nearest ? (tid & (kGroup - 1)) : 0;
        const unsigned gmask = ((1u << kGroup) - 1u) << (lane & ~(kGroup - 1));
        for (int i = gtid / per; i < a.N; i += gthreads / per) {
            bool active;
            double row[6], z, absres;
            lio_point<6>(a, s_pose, nearest != 0, i, sub, gmask, active, row, z, absres);
            if (active) {
                int k = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
#pragma unroll
                    for (int c = r; c < 6; ++c) { acc[k] += row[r] * row[c]; ++k; }
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) acc[21 + r] += row[r] * z;
                acc[27] += 1.0;
                acc[28] += absres;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double v = warp_sum(acc[k]);
            if (lane == 0) s_acc[warp][k] = v;
        }
        __syncthreads();
        for (int q = tid; q < K; q += BLOCK) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += s_acc[w][q];
            a.partials[(size_t)blockIdx.x * K + q] = v;
        }
        const bool leader = grid_arrive(bar, gridDim.x, s_bar);
        if (leader) {
            if (trace && tid == 0 && 2 + 2 * pass_no < kTraceLen) trace[1 + 2 * pass_no] = global_ns();
            load_state_l2(&sm.x, s.state, tid, BLOCK);
            load_state_l2(&sm.xp, s.state_prop, tid, BLOCK);
            leader_reduce<K>(a.partials, gridDim.x, sm, tid);
            if (tid == 0) unpack_sym6(sm.packed, sm.HTH, sm.HTz);
            __syncthreads();
            const bool ok = leader_solve<BLOCK - 32>(sm, s.Pinv, s.prm.sigma, +1.0, first, tid);
            if (tid == 0) {
                LioCtrl c;
                if (first) {
                    c.passes = c.knn_passes = 0;
                    c.rows_total = 0;
                    c.status = 0;
                } else {
                    c.passes = __ldcg(&s.ctrl->passes);
                    c.knn_passes = __ldcg(&s.ctrl->knn_passes);
                    c.rows_total = __ldcg(&s.ctrl->rows_total);
                    c.status = __ldcg(&s.ctrl->status);
                }
                const int n_eff = (int)sm.packed[K - 2];
                c.passes += 1;
                c.knn_passes += nearest ? 1 : 0;
                c.n_eff_last = n_eff;
                c.res_mean_last = sm.packed[K - 1] / (double)n_eff;
                c.rows_total += n_eff;
                state_boxplus(sm.x, sm.sol);                                                   // :1683
                bool converged = false;
                if ((norm3(sm.sol) * 57.3 < s.prm.conv_rot_deg) && (norm3(sm.sol + 3) * 100 < s.prm.conv_pos_cm))
                    converged = true;                                                          // :1688
                int nn = 0, rn = rematch_num;
                if (converged || ((rn == 0) && (iterCount == (T - 2)))) { nn = 1; rn++; }       // :1700-1705
                int stop = 0;
                if (rn >= 2 || (iterCount == T - 1)) stop = 1;                                  // :1708
                if (!ok) { stop = 1; c.status = -5; }
                c.converged_last = converged ? 1 : 0;
                c.iterCount = iterCount + 1;
                c.rematch_num = rn;
                c.nearest_search_en = nn;
                c.stop = stop;
                *s.ctrl = c;
                sm.flags[0] = stop;
                sm.flags[1] = ok ? 1 : 0;
            }
            __syncthreads();
            const bool do_cov = sm.flags[0] && sm.flags[1];
            if (do_cov) leader_cov_update(sm, sm.Gc, tid, BLOCK);                               // :1715
            store_state(s.state, &sm.x, tid, BLOCK, do_cov);
            if (trace && tid == 0 && 2 + 2 * pass_no < kTraceLen) trace[2 + 2 * pass_no] = global_ns();
            __threadfence();
            grid_release(bar, s_bar);
        } else {
            if (!grid_wait(bar, s_bar)) return;
        }
        // every block: pick up the published control through L2
        const int stop = __ldcg(&s.ctrl->stop);
        nearest = __ldcg(&s.ctrl->nearest_search_en);
        rematch_num = __ldcg(&s.ctrl->rematch_num);
        iterCount = __ldcg(&s.ctrl->iterCount);
        first = false;
        ++pass_no;
        if (stop) break;
    }
}

// ---------------------------------------------------------------------------------------
// VIO: ComputeJ (3 levels x up to T passes) in one launch
// ---------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 2) k_vio_update_persistent(VioArgs a, VioSolveArgs s, GridBarrier* bar,
                                                                 unsigned long long* trace) {
    constexpr int NW = BLOCK / 32;
    __shared__ VioPose s_pose;
    __shared__ float s_lat[NW][128];
    __shared__ double s_res[NW][64];
    __shared__ double s_acc[NW][kVioPacked];
    __shared__ unsigned s_bar[2];
    __shared__ LeaderSmem sm;
    __shared__ float s_error;
    constexpr int kErrChunk = 2048;
    __shared__ float s_err[kErrChunk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.Pn <= 0) return;                                     // :969-970 (host also short-circuits)
    int level = 2;
    bool first = true;
    int pass_no = 0;
    if (trace && blockIdx.x == 0 && tid == 0) trace[0] = global_ns();
    for (;;) {
        if (tid == 0) {
            double R[9], p[3];
            for (int i = 0; i < 9; ++i) R[i] = __ldcg(&a.state->rot[i]);
            for (int i = 0; i < 3; ++i) p[i] = __ldcg(&a.state->pos[i]);
            vio_make_pose(a.prm.Rci, a.prm.Pci, R, p, s_pose);
        }
        __syncthreads();
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[k] = 0.0;
        double n_meas = 0.0, skipped = 0.0;
        for (int i = blockIdx.x * NW + warp; i < a.Pn; i += gridDim.x * NW) {
            PatchGeom g;
            const double pos[3] = {a.pos[3 * (size_t)i], a.pos[3 * (size_t)i + 1], a.pos[3 * (size_t)i + 2]};
            vio_patch_geom(a.cam, s_pose, pos, level, a.search_level[i], g);
            if (g.valid) {
                const int W = a.cam.width;
                const unsigned char* base = a.img + (size_t)(g.v_i - 5 * g.scale) * W + (g.u_i - 5 * g.scale);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = lane + 32 * q;
                    if (e < 121) {
                        const int r = e / 11, c = e - r * 11;
                        s_lat[warp][e] = (float)__ldg(base + (size_t)r * g.scale * W + c * g.scale);
                    }
                }
                __syncwarp();
                LatView L{s_lat[warp]};
                const int x = lane >> 2, y0 = (lane & 3) * 2;
                const float* P = a.patch + (size_t)i * 192 + 64 * level;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int y = y0 + q;
                    double row[6], res;
                    vio_pixel(L, g, s_pose, x, y, __ldg(P + x * 8 + y), row, &res);
                    s_res[warp][x * 8 + y] = res;
                    int k = 0;
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
#pragma unroll
                        for (int c = r; c < 6; ++c) { acc[k] += row[r] * row[c]; ++k; }
                    }
#pragma unroll
                    for (int r = 0; r < 6; ++r) acc[21 + r] += row[r] * res;
                }
                __syncwarp();
                if (lane == 0) {
                    float pe = 0.0f;
                    for (int e = 0; e < 64; ++e) pe = (float)((double)pe + s_res[warp][e] * s_res[warp][e]);   // :843
                    a.errors[i] = pe;                                                                          // :851
                    n_meas += 64.0;
                }
                __syncwarp();
            } else if (lane == 0) {
                a.errors[i] = 0.0f;
                skipped += 1.0;
            }
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const double v = warp_sum(acc[k]);
            if (lane == 0) s_acc[warp][k] = v;
        }
        if (lane == 0) { s_acc[warp][27] = n_meas; s_acc[warp][28] = skipped; }
        __syncthreads();
        for (int q = tid; q < kVioPacked; q += BLOCK) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += s_acc[w][q];
            a.partials[(size_t)blockIdx.x * kVioPacked + q] = v;
        }
        const bool leader = grid_arrive(bar, gridDim.x, s_bar);
        if (leader) {
            if (trace && tid == 0 && 2 + 2 * pass_no < kTraceLen) trace[1 + 2 * pass_no] = global_ns();
            load_state_l2(&sm.x, s.state, tid, BLOCK);
            load_state_l2(&sm.xp, s.state_prop, tid, BLOCK);
            if (first) {
                // old_state = *state (:747)
                store_state(s.old_state, &sm.x, tid, BLOCK, true);
            }
            leader_reduce<kVioPacked>(a.partials, gridDim.x, sm, tid);
            if (tid == 0) unpack_sym6(sm.packed, sm.HTH, sm.HTz);
            __syncthreads();
            // One thread forms the exact sequential float sum of the per-patch errors (:852) while the
            // others run the speculative solve.  The errors are first staged into shared memory by the
            // whole block (coalesced L2 reads), so the serial chain costs one FADD latency per patch
            // instead of one L2 round trip per batch.
            float e_run = 0.0f;
            for (int base = 0; base < s.Pn_total; base += kErrChunk) {
                const int nchunk = min(kErrChunk, s.Pn_total - base);
                for (int e = tid; e < nchunk; e += BLOCK) s_err[e] = __ldcg(s.errors + base + e);
                __syncthreads();
                if (tid == BLOCK - 1) {
                    float e = e_run;
                    int i = 0;
                    for (; i + 8 <= nchunk; i += 8) {
                        e = e + s_err[i]; e = e + s_err[i + 1]; e = e + s_err[i + 2]; e = e + s_err[i + 3];
                        e = e + s_err[i + 4]; e = e + s_err[i + 5]; e = e + s_err[i + 6]; e = e + s_err[i + 7];
                    }
                    for (; i < nchunk; ++i) e = e + s_err[i];
                    e_run = e;
                    if (base + kErrChunk >= s.Pn_total) s_error = e;
                }
                if (base + kErrChunk < s.Pn_total) __syncthreads();
            }
            const bool ok = leader_solve<BLOCK - 64>(sm, s.Pinv, s.prm.sigma, -1.0, first, tid);
            if (tid == 0) {
                VioCtrl c;
                if (first) {
                    c.level = 2; c.iteration = 0; c.stop = 0;
                    c.last_error = 1e10f; c.now_error = 1e10f; c.any_solved = 0;
                    for (int l = 0; l < 3; ++l) { c.passes[l] = 0; c.level_error[l] = 1e10f; }
                    c.rows_total = 0; c.skipped_last = 0; c.cov_updated = 0; c.status = 0;
                } else {
                    const int* src = reinterpret_cast<const int*>(s.ctrl);
                    int* dst = reinterpret_cast<int*>(&c);
                    for (int e = 0; e < (int)(sizeof(VioCtrl) / sizeof(int)); ++e) dst[e] = __ldcg(src + e);
                }
                const long long nm = (long long)sm.packed[27];
                const float error = s_error / (float)(unsigned long long)nm;                       // :857
                c.passes[level] += 1;
                c.rows_total += nm;
                c.skipped_last = (int)sm.packed[28];
                bool EKF_end = false;
                int accept = 0;
                if (s.prm.force_all_passes || error <= c.last_error) {                             // :861
                    accept = 1;
                    c.last_error = error;
                    if (!s.prm.force_all_passes && (norm3(sm.sol) * 57.3f < s.prm.conv_rot_deg) &&
                        (norm3(sm.sol + 3) * 100.0f < s.prm.conv_pos_cm))
                        EKF_end = true;                                                            // :883
                    c.any_solved = 1;
                } else {
                    EKF_end = true;                                                                // :890
                }
                if (!ok) { EKF_end = true; c.status = -5; }
                c.iteration += 1;
                int docov = 0, newlevel = 0;
                if (EKF_end || c.iteration >= s.prm.max_iteration) {
                    c.level_error[level] = c.last_error;
                    c.now_error = c.last_error;
                    c.level = level - 1;
                    c.iteration = 0;
                    c.last_error = 1e10f;
                    newlevel = 1;
                    if (c.level < 0 || !ok) {
                        c.stop = 1;
                        if (c.now_error < 1e10f && ok) { docov = 1; c.cov_updated = 1; }           // :978-981
                    }
                }
                *s.ctrl = c;
                sm.flags[0] = accept;
                sm.flags[1] = docov;
                sm.flags[2] = newlevel;
            }
            __syncthreads();
            const int accept = sm.flags[0], docov = sm.flags[1], newlevel = sm.flags[2];
            if (accept) {
                store_state(s.old_state, &sm.x, tid, BLOCK, true);                                 // old_state = *state (:863)
                for (int e = tid; e < kDim * 6; e += BLOCK) s.G_last[e] = sm.Gc[e];
                __syncthreads();
                if (tid == 0) state_boxplus(sm.x, sm.sol);                                         // :879
                __syncthreads();
            } else {
                load_state_l2(&sm.x, s.old_state, tid, BLOCK);                                     // *state = old_state (:890)
                __syncthreads();
            }
            if (docov) {
                if (!accept) {
                    for (int e = tid; e < kDim * 6; e += BLOCK) sm.Gc[e] = __ldcg(s.G_last + e);
                    __syncthreads();
                }
                leader_cov_update(sm, sm.Gc, tid, BLOCK);                                          // :980
            }
            store_state(s.state, &sm.x, tid, BLOCK, true);
            if (newlevel) store_state(s.old_state, &sm.x, tid, BLOCK, true);                       // :747 of the next level
            if (trace && tid == 0 && 2 + 2 * pass_no < kTraceLen) trace[2 + 2 * pass_no] = global_ns();
            __threadfence();
            grid_release(bar, s_bar);
        } else {
            if (!grid_wait(bar, s_bar)) return;
        }
        const int stop = __ldcg(&s.ctrl->stop);
        level = __ldcg(&s.ctrl->level);
        first = false;
        ++pass_no;
        if (stop) break;
    }
}

}  // namespace flb
