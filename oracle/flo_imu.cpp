/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see flo_oracle.h).
 *
 * flo_imu_undistort: CPU restatement of ImuProcess::UndistortPcl
 * (reference src/IMU_Processing.cpp:611-809): IMU forward propagation of
 * state and 18x18 covariance, frame-end prediction, per-point backward motion compensation.
 * The LidarMeasureGroup bookkeeping of :613-652 (which points belong to this update, pcl_beg_time,
 * pcl_end_time) is the caller's; this function starts at :655.
 * PARITY STATUS: "parity unpinned" (the reference has no fixtures for this step either).
 */
#include <cmath>
#include <cstring>
#include <vector>

#include "flo_oracle.h"

namespace {

void m3mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j];
            C[i * 3 + j] = s;
        }
}
void m3vec(const double* A, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
void m3T(const double* A, double* T) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}
void skew(const double* v, double* K) {   // SKEW_SYM_MATRX, include/common_lib.h
    K[0] = 0; K[1] = -v[2]; K[2] = v[1];
    K[3] = v[2]; K[4] = 0; K[5] = -v[0];
    K[6] = -v[1]; K[7] = v[0]; K[8] = 0;
}
/* Exp(ang_vel, dt), include/so3_math.h:30-51 */
void exp_w_dt(const double* w, double dt, double* R) {
    const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (n > 0.0000001) {
        const double ax[3] = {w[0] / n, w[1] / n, w[2] / n};
        double K[9], KK[9];
        skew(ax, K);
        m3mul(K, K, KK);
        const double ang = n * dt, s = std::sin(ang), c = 1.0 - std::cos(ang);
        for (int i = 0; i < 9; ++i) R[i] = R[i] + s * K[i] + c * KK[i];
    }
}

struct Pose6 { double t, acc[3], gyr[3], vel[3], pos[3], rot[9]; };

}  // namespace

extern "C" int flo_imu_undistort(const flo_imu_params* prm, flo_imu_carry* carry, const flo_imu_sample* v_imu, int n_imu,
                                 double pcl_beg_time, double pcl_end_time, flo_state18* x, float* pts_xyz,
                                 const float* offset_ms, int n_points, int* n_poses_out, double* poses_out) {
    if (n_imu < 1) return -1;
    const double imu_end_time = v_imu[n_imu - 1].t;
    std::vector<Pose6> P;
    {
        Pose6 p0;
        p0.t = 0.0;
        std::memcpy(p0.acc, carry->acc_s_last, 24);
        std::memcpy(p0.gyr, carry->angvel_last, 24);
        std::memcpy(p0.vel, x->vel, 24);
        std::memcpy(p0.pos, x->pos, 24);
        std::memcpy(p0.rot, x->rot, 72);
        P.push_back(p0);                                                         /* :658 */
    }
    double acc_imu[3], angvel_avr[3], acc_avr[3], vel_imu[3], pos_imu[3], R_imu[9];
    std::memcpy(acc_imu, carry->acc_s_last, 24);
    std::memcpy(angvel_avr, carry->angvel_last, 24);
    std::memcpy(vel_imu, x->vel, 24);
    std::memcpy(pos_imu, x->pos, 24);
    std::memcpy(R_imu, x->rot, 72);
    double dt = 0.0;
    std::vector<double> F(324), W(324), T(324);
    for (int it = 0; it + 1 < n_imu; ++it) {                                     /* :666 */
        const flo_imu_sample& head = v_imu[it];
        const flo_imu_sample& tail = v_imu[it + 1];
        if (tail.t < carry->last_lidar_end_time) continue;                       /* :671 */
        for (int k = 0; k < 3; ++k) {
            angvel_avr[k] = 0.5 * (head.gyr[k] + tail.gyr[k]);                   /* :673-681 */
            acc_avr[k] = 0.5 * (head.acc[k] + tail.acc[k]);
        }
        for (int k = 0; k < 3; ++k) {
            angvel_avr[k] -= x->bg[k];                                           /* :687 */
            acc_avr[k] = acc_avr[k] * prm->G_m_s2 / prm->mean_acc_norm - x->ba[k]; /* :688 */
        }
        if (head.t < carry->last_lidar_end_time) dt = tail.t - carry->last_lidar_end_time;   /* :690-697 */
        else dt = tail.t - head.t;
        double Exp_f[9], Exp_m[9], askew[9], Ra[9];
        exp_w_dt(angvel_avr, dt, Exp_f);                                         /* :701 */
        exp_w_dt(angvel_avr, -dt, Exp_m);                                        /* :707 */
        skew(acc_avr, askew);
        std::fill(F.begin(), F.end(), 0.0);
        std::fill(W.begin(), W.end(), 0.0);
        for (int i = 0; i < 18; ++i) F[i * 18 + i] = 1.0;
        m3mul(R_imu, askew, Ra);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                F[i * 18 + j] = Exp_m[i * 3 + j];                                /* (0,0) */
                F[i * 18 + 9 + j] = (i == j) ? -dt : 0.0;                        /* (0,9)  = -I dt */
                F[(3 + i) * 18 + 6 + j] = (i == j) ? dt : 0.0;                   /* (3,6)  =  I dt */
                F[(6 + i) * 18 + j] = -Ra[i * 3 + j] * dt;                       /* (6,0)  = -R [a]x dt */
                F[(6 + i) * 18 + 12 + j] = -R_imu[i * 3 + j] * dt;               /* (6,12) = -R dt */
                F[(6 + i) * 18 + 15 + j] = (i == j) ? dt : 0.0;                  /* (6,15) =  I dt */
            }
        for (int i = 0; i < 3; ++i) {
            W[i * 18 + i] = prm->cov_gyr[i] * dt * dt;                           /* :715 */
            W[(9 + i) * 18 + 9 + i] = prm->cov_bias_gyr[i] * dt * dt;            /* :717 */
            W[(12 + i) * 18 + 12 + i] = prm->cov_bias_acc[i] * dt * dt;          /* :718 */
        }
        {   /* cov_w(6,6) = R diag(cov_acc) R^T dt^2, :716 (left to right) */
            double RD[9], Rt[9], RDRt[9];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) RD[i * 3 + j] = R_imu[i * 3 + j] * prm->cov_acc[j];
            m3T(R_imu, Rt);
            m3mul(RD, Rt, RDRt);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) W[(6 + i) * 18 + 6 + j] = RDRt[i * 3 + j] * dt * dt;
        }
        /* cov = F cov F^T + cov_w, :720 */
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) {
                double s = 0.0;
                for (int k = 0; k < 18; ++k) s += F[i * 18 + k] * x->cov[k * 18 + j];
                T[i * 18 + j] = s;
            }
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) {
                double s = 0.0;
                for (int k = 0; k < 18; ++k) s += T[i * 18 + k] * F[j * 18 + k];
                x->cov[i * 18 + j] = s + W[i * 18 + j];
            }
        double Rn[9], Ra_v[3];
        m3mul(R_imu, Exp_f, Rn);                                                 /* :723 */
        std::memcpy(R_imu, Rn, 72);
        m3vec(R_imu, acc_avr, Ra_v);
        for (int k = 0; k < 3; ++k) acc_imu[k] = Ra_v[k] + x->grav[k];           /* :726 */
        for (int k = 0; k < 3; ++k) pos_imu[k] = pos_imu[k] + vel_imu[k] * dt + 0.5 * acc_imu[k] * dt * dt;   /* :729 */
        for (int k = 0; k < 3; ++k) vel_imu[k] = vel_imu[k] + acc_imu[k] * dt;   /* :732 */
        std::memcpy(carry->angvel_last, angvel_avr, 24);                         /* :735-736 */
        std::memcpy(carry->acc_s_last, acc_imu, 24);
        Pose6 p;
        p.t = tail.t - pcl_beg_time;                                             /* :737 */
        std::memcpy(p.acc, acc_imu, 24);
        std::memcpy(p.gyr, angvel_avr, 24);
        std::memcpy(p.vel, vel_imu, 24);
        std::memcpy(p.pos, pos_imu, 24);
        std::memcpy(p.rot, R_imu, 72);
        P.push_back(p);
    }
    /* frame-end prediction, :743-758 */
    {
        double note;
        if (imu_end_time > pcl_beg_time) {
            note = pcl_end_time > imu_end_time ? 1.0 : -1.0;
            dt = note * (pcl_end_time - imu_end_time);
        } else {
            note = pcl_end_time > pcl_beg_time ? 1.0 : -1.0;
            dt = note * (pcl_end_time - pcl_beg_time);
        }
        double w[3] = {note * angvel_avr[0], note * angvel_avr[1], note * angvel_avr[2]}, E[9], Rn[9];
        exp_w_dt(w, dt, E);
        m3mul(R_imu, E, Rn);
        for (int k = 0; k < 3; ++k) {
            x->vel[k] = vel_imu[k] + note * acc_imu[k] * dt;
            x->pos[k] = pos_imu[k] + note * vel_imu[k] * dt + note * 0.5 * acc_imu[k] * dt * dt;
        }
        std::memcpy(x->rot, Rn, 72);
    }
    carry->last_lidar_end_time = pcl_end_time;                                   /* :761 */
    if (n_poses_out) *n_poses_out = (int)P.size();
    if (poses_out)
        for (size_t k = 0; k < P.size(); ++k) std::memcpy(poses_out + 22 * k, &P[k], sizeof(Pose6));
    double RLIt[9], rot_end_T[9], extR_Ri[9], exrR_extT[3];
    m3T(prm->R_LI, RLIt);
    m3T(x->rot, rot_end_T);
    m3mul(RLIt, rot_end_T, extR_Ri);                                             /* :763 */
    m3vec(RLIt, prm->t_LI, exrR_extT);                                           /* :764 */
    if (n_points < 1) return 0;                                                  /* :776 */
    /* backward undistortion, :778-808.  Restated loop for loop: after the `break` at the first point (:807) the
     * outer loop goes on to the earlier IMU poses and tests that same point again, so a first point that is
     * later than an earlier pose's offset is compensated once more per such pose (on its already compensated,
     * float-rounded coordinates). */
    int ip = n_points - 1;
    for (int kp = (int)P.size() - 1; kp > 0; --kp) {
        const Pose6& head = P[kp - 1];
        for (; (double)offset_ms[ip] / double(1000) > head.t; --ip) {
            dt = (double)offset_ms[ip] / double(1000) - head.t;
            double E[9], R_i[9], T_ei[3], P_i[3], a[3], b[3], c[3];
            exp_w_dt(head.gyr, dt, E);
            m3mul(head.rot, E, R_i);                                             /* :796 */
            for (int k = 0; k < 3; ++k) T_ei[k] = head.pos[k] + head.vel[k] * dt + 0.5 * head.acc[k] * dt * dt - x->pos[k];   /* :797 */
            for (int k = 0; k < 3; ++k) P_i[k] = (double)pts_xyz[3 * ip + k];
            m3vec(prm->R_LI, P_i, a);
            for (int k = 0; k < 3; ++k) a[k] += prm->t_LI[k];
            m3vec(R_i, a, b);
            for (int k = 0; k < 3; ++k) b[k] += T_ei[k];
            m3vec(extR_Ri, b, c);
            for (int k = 0; k < 3; ++k) pts_xyz[3 * ip + k] = (float)(c[k] - exrR_extT[k]);   /* :800-805 */
            if (ip == 0) break;                                                  /* :807 */
        }
    }
    return 0;
}
