// TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see flo_oracle.h).  "parity unpinned".
//
// CPU restatement of the IKFoM-typed estimator path (dead code at the reference commit, but named
// by north_star / SURVEY.md section 8 row a8):
//   esekfom::esekf<state_ikfom,12,input_ikfom>::update_iterated_dyn_share_modified
//       include/IKFoM_toolkit/esekfom/esekfom.hpp:1619-1928
//   state_ikfom = pos, rot(SO3), offset_R_L_I(SO3), offset_T_L_I, vel, bg, ba, grav(S2, |g| = 9.8090)
//       include/use-ikfom.hpp:12-21            (DOF 23, order as listed)
//   MTK::SO3 boxplus / boxminus / exp / log    include/IKFoM_toolkit/mtk/types/SOn.hpp:233-299
//   MTK::S2 boxplus / boxminus / Bx / Nx_yy / Mx   include/IKFoM_toolkit/mtk/types/S2.hpp:97-280
//   MTK::A_matrix, cos_sinc_sqrt, exp, log     include/IKFoM_toolkit/mtk/src/mtkmath.hpp:141-297
// driven by the measurement model h_share_model (src/laserMapping.cpp:960-1094), for which this file
// calls flo_lio_pass (12-column IKFoM row layout) with R = rot.toRotationMatrix() etc.  (The reference
// rotates with the quaternion directly; the two differ by ~1e-16.)
#include "flo_oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr int N = 23;                 // DOF of state_ikfom
constexpr double kTol = 1e-11;        // MTK::tolerance<double>()
constexpr double kLen = 98090.0 / 10000.0;   // S2<double, 98090, 10000, 1>::length

// ---- small helpers ----------------------------------------------------------------------
inline void hat(const double* v, double* K) {
    K[0] = 0; K[1] = -v[2]; K[2] = v[1];
    K[3] = v[2]; K[4] = 0; K[5] = -v[0];
    K[6] = -v[1]; K[7] = v[0]; K[8] = 0;
}
inline void mm(const double* A, int ra, int ca, const double* B, int cb, double* C) {   // C(ra x cb) = A(ra x ca) B(ca x cb)
    for (int i = 0; i < ra; ++i)
        for (int j = 0; j < cb; ++j) {
            double s = 0;
            for (int k = 0; k < ca; ++k) s += A[i * ca + k] * B[k * cb + j];
            C[i * cb + j] = s;
        }
}
inline void transpose(const double* A, int r, int c, double* T) {
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) T[j * r + i] = A[i * c + j];
}

// mtkmath.hpp:141-171
void cos_sinc_sqrt(double x2, double* c, double* sc) {
    const double taylor_0 = 2.220446049250313e-16;
    const double taylor_2 = std::sqrt(taylor_0);
    const double taylor_n = std::sqrt(taylor_2);
    if (x2 >= taylor_n) {
        const double x = std::sqrt(x2);
        *c = std::cos(x);
        *sc = std::sin(x) / x;
        return;
    }
    static const double inv[] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
    double cosi = 1., sinc = 1;
    double term = -1 / 2. * x2;
    for (int i = 0; i < 3; ++i) {
        cosi += term;
        term *= inv[2 * i];
        sinc += term;
        term *= -inv[2 * i + 1] * x2;
    }
    *c = cosi;
    *sc = sinc;
}

// Quaternions are stored (x, y, z, w) like Eigen's coeffs().
// MTK::exp<scalar,3> (mtkmath.hpp:249-256) into a quaternion: SO3::exp(vec, scale) (SOn.hpp:282-286)
void quat_exp(const double* v, double scale, double* q) {
    const double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double c, sc;
    cos_sinc_sqrt(scale * scale * n2, &c, &sc);
    const double mult = sc * scale;
    q[0] = mult * v[0]; q[1] = mult * v[1]; q[2] = mult * v[2];
    q[3] = c;
}
void quat_mul(const double* a, const double* b, double* o) {   // Eigen: a * b
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
void quat_to_R(const double* q, double* R) {   // Eigen::Quaternion::toRotationMatrix
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
                 tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// SO3::log (SOn.hpp:291-295) -> MTK::log<scalar,3>(res, w, vec, 2, true) (mtkmath.hpp:269-289)
void quat_log(const double* q, double* out) {
    double nv = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (nv < kTol) nv = kTol;
    const double s = 2.0 / nv * std::atan(nv / q[3]);
    for (int i = 0; i < 3; ++i) out[i] = s * q[i];
}
// MTK::A_matrix, mtkmath.hpp:235-247
void A_matrix(const double* v, double* A) {
    const double sq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double norm = std::sqrt(sq);
    for (int i = 0; i < 9; ++i) A[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (norm < kTol) return;
    double K[9], KK[9];
    hat(v, K);
    mm(K, 3, 3, K, 3, KK);
    const double a = (1 - std::cos(norm)) / sq, b = (1 - std::sin(norm) / norm) / sq;
    for (int i = 0; i < 9; ++i) A[i] = A[i] + a * K[i] + b * KK[i];
}

// ---- S2 (S2_typ = 1, length 9.8090), S2.hpp ---------------------------------------------------
void S2_Bx(const double* vec, double* Bx /*3x2*/) {   // S2.hpp:206-231 (S2_typ == 1 branch)
    if (vec[0] + kLen > kTol) {
        const double d = kLen + vec[0];
        Bx[0] = -vec[1];                    Bx[1] = -vec[2];
        Bx[2] = kLen - vec[1] * vec[1] / d; Bx[3] = -vec[2] * vec[1] / d;
        Bx[4] = -vec[2] * vec[1] / d;       Bx[5] = kLen - vec[2] * vec[2] / d;
        for (int i = 0; i < 6; ++i) Bx[i] /= kLen;
    } else {
        for (int i = 0; i < 6; ++i) Bx[i] = 0;
        Bx[1 * 2 + 1] = -1;
        Bx[2 * 2 + 0] = 1;
    }
}
void S2_boxplus(double* vec, const double* delta) {   // S2.hpp:112-118
    double Bx[6], Bu[3], q[4], R[9], o[3];
    S2_Bx(vec, Bx);
    mm(Bx, 3, 2, delta, 1, Bu);
    quat_exp(Bu, 0.5, q);
    quat_to_R(q, R);
    mm(R, 3, 3, vec, 1, o);
    std::memcpy(vec, o, sizeof(o));
}
void S2_boxminus(const double* vec, const double* other, double* res) {   // S2.hpp:120-144
    double K[9], t[3];
    hat(vec, K);
    mm(K, 3, 3, other, 1, t);
    const double v_sin = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    const double v_cos = vec[0] * other[0] + vec[1] * other[1] + vec[2] * other[2];
    const double theta = std::atan2(v_sin, v_cos);
    if (v_sin < kTol) {
        if (std::fabs(theta) > kTol) { res[0] = 3.1415926; res[1] = 0; }
        else { res[0] = 0; res[1] = 0; }
    } else {
        double Bx[6], BxT[6], Ko[9], u[3], r2[2];
        S2_Bx(other, Bx);
        transpose(Bx, 3, 2, BxT);
        hat(other, Ko);
        mm(Ko, 3, 3, vec, 1, u);
        mm(BxT, 2, 3, u, 1, r2);
        res[0] = theta / v_sin * r2[0];
        res[1] = theta / v_sin * r2[1];
    }
}
void S2_Nx_yy(const double* vec, double* Nx /*2x3*/) {   // S2.hpp:262-267
    double Bx[6], BxT[6], K[9];
    S2_Bx(vec, Bx);
    transpose(Bx, 3, 2, BxT);
    hat(vec, K);
    mm(BxT, 2, 3, K, 3, Nx);
    for (int i = 0; i < 6; ++i) Nx[i] = 1 / kLen / kLen * Nx[i];
}
void S2_Mx(const double* vec, const double* delta, double* Mx /*3x2*/) {   // S2.hpp:269-283
    double Bx[6], K[9];
    S2_Bx(vec, Bx);
    hat(vec, K);
    if (std::sqrt(delta[0] * delta[0] + delta[1] * delta[1]) < kTol) {
        mm(K, 3, 3, Bx, 2, Mx);
        for (int i = 0; i < 6; ++i) Mx[i] = -Mx[i];
    } else {
        // NB: the reference passes scalar(1/2) -- INTEGER division, i.e. scale 0 -- to MTK::exp here
        // (S2.hpp:280), so exp_delta is the identity rotation.  Mirrored.
        double Bu[3], q[4], R[9], A[9], AT[9], t1[9], t2[9];
        mm(Bx, 3, 2, delta, 1, Bu);
        quat_exp(Bu, 0.0, q);
        quat_to_R(q, R);
        A_matrix(Bu, A);
        transpose(A, 3, 3, AT);
        mm(R, 3, 3, K, 3, t1);
        mm(t1, 3, 3, AT, 3, t2);
        mm(t2, 3, 3, Bx, 2, Mx);
        for (int i = 0; i < 6; ++i) Mx[i] = -Mx[i];
    }
}

// Apply a k x k matrix T to rows [idx, idx+k) of an N x N matrix: M[idx.., c] = T * M[idx.., c]
template <int K>
void rows_apply(double* M, int idx, const double* T, const double* src) {
    for (int c = 0; c < N; ++c) {
        double v[K];
        for (int i = 0; i < K; ++i) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += T[i * K + k] * src[(idx + k) * N + c];
            v[i] = s;
        }
        for (int i = 0; i < K; ++i) M[(idx + i) * N + c] = v[i];
    }
}
// M[r, idx..] = M[r, idx..] * T^T
template <int K>
void cols_apply_T(double* M, int idx, const double* T) {
    for (int r = 0; r < N; ++r) {
        double v[K];
        for (int j = 0; j < K; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += M[r * N + idx + k] * T[j * K + k];
            v[j] = s;
        }
        for (int j = 0; j < K; ++j) M[r * N + idx + j] = v[j];
    }
}

}  // namespace

extern "C" {

// state_ikfom::boxplus (build_manifold.hpp:192-196), sub-manifold order of use-ikfom.hpp:12-21
void flo_ikfom_boxplus(flo_state_ikfom* x, const double* d) {
    for (int i = 0; i < 3; ++i) x->pos[i] += d[i];
    double q[4], o[4];
    quat_exp(d + 3, 0.5, q);
    quat_mul(x->rot, q, o);
    std::memcpy(x->rot, o, sizeof(o));
    quat_exp(d + 6, 0.5, q);
    quat_mul(x->offset_R_L_I, q, o);
    std::memcpy(x->offset_R_L_I, o, sizeof(o));
    for (int i = 0; i < 3; ++i) {
        x->offset_T_L_I[i] += d[9 + i];
        x->vel[i] += d[12 + i];
        x->bg[i] += d[15 + i];
        x->ba[i] += d[18 + i];
    }
    S2_boxplus(x->grav, d + 21);
}

// res = a [-] b  (state_ikfom::boxminus, build_manifold.hpp:198-202)
void flo_ikfom_boxminus(const flo_state_ikfom* a, const flo_state_ikfom* b, double* res) {
    for (int i = 0; i < 3; ++i) res[i] = a->pos[i] - b->pos[i];
    double c[4], q[4];
    c[0] = -b->rot[0]; c[1] = -b->rot[1]; c[2] = -b->rot[2]; c[3] = b->rot[3];
    quat_mul(c, a->rot, q);
    quat_log(q, res + 3);
    c[0] = -b->offset_R_L_I[0]; c[1] = -b->offset_R_L_I[1]; c[2] = -b->offset_R_L_I[2]; c[3] = b->offset_R_L_I[3];
    quat_mul(c, a->offset_R_L_I, q);
    quat_log(q, res + 6);
    for (int i = 0; i < 3; ++i) {
        res[9 + i] = a->offset_T_L_I[i] - b->offset_T_L_I[i];
        res[12 + i] = a->vel[i] - b->vel[i];
        res[15 + i] = a->bg[i] - b->bg[i];
        res[18 + i] = a->ba[i] - b->ba[i];
    }
    S2_boxminus(a->grav, b->grav, res + 21);
}

void flo_quat_to_R(const double* q, double* R) { quat_to_R(q, R); }

// update_iterated_dyn_share_modified, esekfom.hpp:1619-1928, with h_dyn_share = h_share_model.
int flo_ikfom_update(flo_lio* L, const flo_ikfom_params* prm, flo_state_ikfom* x, flo_ikfom_report* rep) {
    const double Rcov = prm->laser_point_cov;
    const int maximum_iter = prm->max_iteration;
    flo_ikfom_report r = {};
    bool converge = true;                                     // dyn_share.converge, :1623
    int t = 0;
    const flo_state_ikfom x_prop = *x;                        // :1625
    std::vector<double> P_prop(x->P, x->P + N * N);           // :1626
    flo_lio_reset(L);
    double dx_new[N];
    for (int i = -1; i < maximum_iter; i++) {                 // :1633
        // ---- h_share_model (src/laserMapping.cpp:960-1094) at the current state
        flo_lio_params lp = {};
        quat_to_R(x->offset_R_L_I, lp.R_LI);
        std::memcpy(lp.t_LI, x->offset_T_L_I, sizeof(lp.t_LI));
        lp.laser_point_cov = Rcov;
        lp.nthreads = prm->nthreads;
        double Rm[9];
        quat_to_R(x->rot, Rm);
        double HTH[144], HTh[12], total = 0.0;
        const int n = flo_lio_pass(L, &lp, Rm, x->pos, converge ? 1 : 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, HTH, HTh, &total);
        r.passes++;
        r.knn_passes += converge ? 1 : 0;
        r.n_eff_last = n;
        r.rows_total += n;
        r.res_mean_last = total / n;
        double dx[N];
        flo_ikfom_boxminus(x, &x_prop, dx);                   // :1646
        std::memcpy(dx_new, dx, sizeof(dx));
        double* P = x->P;
        std::memcpy(P, P_prop.data(), sizeof(double) * N * N);   // :1654
        for (int idx : {3, 6}) {                              // SO3 states, :1658-1674
            double A[9], AT[9], v[3];
            A_matrix(dx + idx, A);
            transpose(A, 3, 3, AT);
            mm(AT, 3, 3, dx_new + idx, 1, v);
            std::memcpy(dx_new + idx, v, sizeof(v));
            rows_apply<3>(P, idx, AT, P);
            cols_apply_T<3>(P, idx, AT);
        }
        {                                                     // S2 state (grav, idx 21), :1676-1699
            const int idx = 21;
            double Nx[6], Mx[6], T2[4], v[2];
            S2_Nx_yy(x->grav, Nx);
            S2_Mx(x_prop.grav, dx + idx, Mx);
            mm(Nx, 2, 3, Mx, 2, T2);
            mm(T2, 2, 2, dx_new + idx, 1, v);
            std::memcpy(dx_new + idx, v, sizeof(v));
            rows_apply<2>(P, idx, T2, P);
            cols_apply_T<2>(P, idx, T2);
        }
        // information form, :1779-1806 (n >= 23).  For n < 23 the reference switches to the algebraically
        // identical Kalman-gain form (:1712-1741); the oracle uses the information form for every n.
        std::vector<double> Ps(N * N), Ptemp(N * N), Pinv(N * N);
        for (int k = 0; k < N * N; ++k) Ps[k] = P[k] / Rcov;
        if (flo_inverse(Ps.data(), N, Ptemp.data()) != 0) return -5;
        for (int a = 0; a < 12; ++a)
            for (int b = 0; b < 12; ++b) Ptemp[a * N + b] += HTH[a * 12 + b];
        if (flo_inverse(Ptemp.data(), N, Pinv.data()) != 0) return -5;
        double K_h[N], K_x[N * N] = {0};
        for (int a = 0; a < N; ++a) {
            double s = 0;
            for (int k = 0; k < 12; ++k) s += Pinv[a * N + k] * HTh[k];
            K_h[a] = s;
            for (int b = 0; b < 12; ++b) {
                double s2 = 0;
                for (int k = 0; k < 12; ++k) s2 += Pinv[a * N + k] * HTH[k * 12 + b];
                K_x[a * N + b] = s2;
            }
        }
        double dx_[N];                                        // :1812
        for (int a = 0; a < N; ++a) {
            double s = K_h[a];
            for (int b = 0; b < N; ++b) s += (K_x[a * N + b] - (a == b ? 1.0 : 0.0)) * dx_new[b];
            dx_[a] = s;
        }
        flo_ikfom_boxplus(x, dx_);                            // :1814
        converge = true;
        for (int a = 0; a < N; ++a)
            if (std::fabs(dx_[a]) > prm->limit[a]) { converge = false; break; }   // :1816-1823
        if (converge) t++;
        if (!t && i == maximum_iter - 2) converge = true;     // :1826-1829
        r.converged_last = converge ? 1 : 0;
        if (t > 1 || i == maximum_iter - 1) {                 // :1831
            std::vector<double> Lm(P, P + N * N);             // L_ = P_
            for (int idx : {3, 6}) {                          // :1836-1859
                double A[9], AT[9];
                A_matrix(dx_ + idx, A);
                transpose(A, 3, 3, AT);
                rows_apply<3>(Lm.data(), idx, AT, P);         // L_.block(idx, c) = A^T * P_.block(idx, c)
                for (int c = 0; c < 12; ++c) {                // K_x rows
                    double v[3];
                    for (int a = 0; a < 3; ++a) {
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += AT[a * 3 + k] * K_x[(idx + k) * N + c];
                        v[a] = s;
                    }
                    for (int a = 0; a < 3; ++a) K_x[(idx + a) * N + c] = v[a];
                }
                cols_apply_T<3>(Lm.data(), idx, AT);
                cols_apply_T<3>(P, idx, AT);
            }
            {                                                 // :1861-1893
                const int idx = 21;
                double Nx[6], Mx[6], T2[4];
                S2_Nx_yy(x->grav, Nx);
                S2_Mx(x_prop.grav, dx_ + idx, Mx);
                mm(Nx, 2, 3, Mx, 2, T2);
                rows_apply<2>(Lm.data(), idx, T2, P);
                for (int c = 0; c < 12; ++c) {
                    double v[2];
                    for (int a = 0; a < 2; ++a) v[a] = T2[a * 2] * K_x[idx * N + c] + T2[a * 2 + 1] * K_x[(idx + 1) * N + c];
                    for (int a = 0; a < 2; ++a) K_x[(idx + a) * N + c] = v[a];
                }
                cols_apply_T<2>(Lm.data(), idx, T2);
                cols_apply_T<2>(P, idx, T2);
            }
            // P_ = L_ - K_x[:, :12] * P_[:12, :]              :1918
            std::vector<double> Pn(N * N);
            for (int a = 0; a < N; ++a)
                for (int b = 0; b < N; ++b) {
                    double s = 0;
                    for (int k = 0; k < 12; ++k) s += K_x[a * N + k] * P[k * N + b];
                    Pn[a * N + b] = Lm[a * N + b] - s;
                }
            std::memcpy(P, Pn.data(), sizeof(double) * N * N);
            if (rep) *rep = r;
            return 0;
        }
    }
    if (rep) *rep = r;
    return 0;
}

}  // extern "C"
