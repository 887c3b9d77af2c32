// fastlivo_b200_adapters.hpp -- header-only C++ adapters from the reference's Eigen / PCL /
// OpenCV-typed call sites onto the C ABI (include/fastlivo_b200.h).
//
// The adapters are templates over the reference's own types so that this header needs NO
// Eigen / PCL / OpenCV include of its own: it compiles inside the reference tree against the
// real types and inside tests/adapters against small mock types.
//
//   flb::lio_update(h, prm, state, state_propagat)      replaces the loop body
//        src/laserMapping.cpp:1504-1733          (live 18-DoF path, StatesGroup)
//   flb::h_share_model(h, prm, s, ekfom_data, ...)      same contract as
//        void h_share_model(state_ikfom&, esekfom::dyn_share_datastruct<double>&)
//        src/laserMapping.cpp:960-1094           (IKFoM-typed callback; rows from the GPU)
//   flb::compute_j(h, prm, img, sub_sparse_map, state, state_propagat)   replaces
//        LidarSelector::ComputeJ(cv::Mat)  src/lidar_selection.cpp:967-983
//   flb::undistort_pcl(h, prm, carry, v_imu, pcl_beg_time, pcl_end_time, state_inout, pcl_out)   replaces
//        ImuProcess::UndistortPcl from src/IMU_Processing.cpp:655 on (propagation + undistortion)
//
// Type requirements (all met by the reference's types):
//   Mat3  : operator()(i,j) read/write (Eigen::Matrix3d)          Vec3 : operator()(i) / [i]
//   Cov   : operator()(i,j), 18x18 (Eigen::Matrix<double,18,18>)
//   StatesGroup : .rot_end .pos_end .vel_end .bias_g .bias_a .gravity .cov  (include/common_lib.h:296-381)
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fastlivo_b200.h"

namespace flb {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(flb_handle* h, int rc) {
    if (rc != FLB_OK) throw Error(rc, flb_last_error(h));
}

// ---- StatesGroup <-> flb_state18 (Eigen is column-major; the ABI is row-major) ------------
template <class StatesGroup>
inline void to_abi(const StatesGroup& s, flb_state18& o) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) o.rot[3 * i + j] = s.rot_end(i, j);
        o.pos[i] = s.pos_end(i);
        o.vel[i] = s.vel_end(i);
        o.bg[i] = s.bias_g(i);
        o.ba[i] = s.bias_a(i);
        o.grav[i] = s.gravity(i);
    }
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) o.cov[18 * i + j] = s.cov(i, j);
}
template <class StatesGroup>
inline void from_abi(const flb_state18& o, StatesGroup& s) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) s.rot_end(i, j) = o.rot[3 * i + j];
        s.pos_end(i) = o.pos[i];
        s.vel_end(i) = o.vel[i];
        s.bias_g(i) = o.bg[i];
        s.bias_a(i) = o.ba[i];
        s.gravity(i) = o.grav[i];
    }
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) s.cov(i, j) = o.cov[18 * i + j];
}

template <class Mat3, class Vec3>
inline flb_lio_params make_lio_params(const Mat3& Lidar_rot_to_IMU, const Vec3& Lidar_offset_to_IMU, double laser_point_cov,
                                      int max_iteration) {
    flb_lio_params p{};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) p.R_LI[3 * i + j] = Lidar_rot_to_IMU(i, j);
        p.t_LI[i] = Lidar_offset_to_IMU(i);
    }
    p.laser_point_cov = laser_point_cov;
    p.max_iteration = max_iteration;
    p.conv_rot_deg = 0.01;    // src/laserMapping.cpp:1688
    p.conv_pos_cm = 0.015;
    return p;
}

// ---- scan / map upload from PCL clouds ---------------------------------------------------
// PointCloud: .points (vector of pcl::PointXYZINormal, 48 bytes, x,y,z first) -- stride 12 floats.
template <class PointCloud>
inline void upload_scan(flb_handle* h, const PointCloud& feats_down_body) {
    static_assert(sizeof(feats_down_body.points[0]) % sizeof(float) == 0, "point type must be float-packed");
    const int stride = (int)(sizeof(feats_down_body.points[0]) / sizeof(float));
    check(h, flb_scan_upload(h, feats_down_body.points.empty() ? nullptr : &feats_down_body.points[0].x,
                             (int)feats_down_body.points.size(), stride));
}
// PointVector = ikd-Tree's flattened storage (KD_TREE::flatten, include/ikd-Tree/ikd_Tree.cpp:1247-1273)
template <class PointVector>
inline void upload_map(flb_handle* h, const PointVector& map_points) {
    const int stride = (int)(sizeof(map_points[0]) / sizeof(float));
    check(h, flb_map_upload(h, &map_points[0].x, (int)map_points.size(), stride));
}

// ---- live LIO: the whole iterated update (src/laserMapping.cpp:1504-1733) -------------------
// On return `state` holds the updated state and covariance; the report carries effct_feat_num,
// res_mean_last and the pass counts the reference prints (:1784-1805).
template <class StatesGroup>
inline flb_lio_report lio_update(flb_handle* h, const flb_lio_params& prm, StatesGroup& state, const StatesGroup& state_propagat) {
    flb_state18 x, xp;
    to_abi(state, x);
    to_abi(state_propagat, xp);
    flb_lio_report rep{};
    check(h, flb_lio_update(h, &prm, &x, &xp, &rep));
    from_abi(x, state);
    return rep;
}

// ---- IKFoM-typed measurement callback (src/laserMapping.cpp:960-1094) ------------------------
// StateIkfom: .pos (vect3), .rot / .offset_R_L_I (MTK SO3, Eigen quaternion: toRotationMatrix()),
//             .offset_T_L_I (vect3)                     (include/use-ikfom.hpp:12-21)
// DynShare  : esekfom::dyn_share_datastruct<double>: .valid .converge .h_x (MatrixXd) .h (VectorXd)
//             (include/IKFoM_toolkit/esekfom/esekfom.hpp:79-89)
// The scan must have been uploaded (upload_scan) before kf.update_iterated_dyn_share_modified().
// Side effects mirrored for the caller: effct_feat_num, res_mean_last, and the scan indices of the
// effective rows (laserCloudOri order) for publish_effect_world (:871-885).
template <class StateIkfom, class DynShare>
inline void h_share_model(flb_handle* h, double laser_point_cov, StateIkfom& s, DynShare& ekfom_data, int& effct_feat_num,
                          double& res_mean_last, std::vector<int>* effective_scan_idx = nullptr) {
    flb_lio_params prm{};
    const auto R_LI = s.offset_R_L_I.toRotationMatrix();
    const auto R = s.rot.toRotationMatrix();
    double Rm[9], pv[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { prm.R_LI[3 * i + j] = R_LI(i, j); Rm[3 * i + j] = R(i, j); }
        prm.t_LI[i] = s.offset_T_L_I(i);
        pv[i] = s.pos(i);
    }
    prm.laser_point_cov = laser_point_cov;
    flb_normal_eq eq{};
    check(h, flb_lio_pass(h, &prm, Rm, pv, ekfom_data.converge ? 1 : 0, 12, &eq));   // :994 converge => redo kNN
    const int n = eq.n_eff;
    std::vector<double> rows((size_t)n * 12 + 1), meas((size_t)n + 1);
    std::vector<int> idx((size_t)n + 1);
    int n_rows = 0;
    check(h, flb_lio_export(h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, rows.data(), meas.data(), idx.data(), &n_rows));
    effct_feat_num = n_rows;                                      // :1039-1052
    res_mean_last = eq.sum_abs_res / n_rows;                      // :1054
    ekfom_data.h_x.resize(n_rows, 12);                            // :1059  (MatrixXd::Zero(effct_feat_num, 12))
    ekfom_data.h.resize(n_rows);                                  // :1060
    for (int r = 0; r < n_rows; ++r) {
        for (int c = 0; c < 12; ++c) ekfom_data.h_x(r, c) = rows[(size_t)r * 12 + c];   // :1082
        ekfom_data.h(r) = meas[r];                                                      // :1088
    }
    if (effective_scan_idx) effective_scan_idx->assign(idx.begin(), idx.begin() + n_rows);
}

// ---- VIO: ComputeJ (src/lidar_selection.cpp:967-983) -------------------------------------------
// Image : cv::Mat-like: .data (uint8*), .cols, .rows, .step (bytes per row)
// SubMap: SubSparseMap (include/common_lib.h:263-292): .voxel_points[i]->pos_ (Vector3d, may be nullptr),
//         .patch[i] (vector<float>, 3*64), .search_levels[i], .errors[i]
// Patches with a null voxel point are dropped (the reference `continue`s over them, :798).
template <class Mat3, class Vec3>
inline flb_vio_params make_vio_params(const Mat3& Rcl, const Vec3& Pcl, const Mat3& Lidar_rot_to_IMU,
                                      const Vec3& Lidar_offset_to_IMU, double img_point_cov, int max_iteration) {
    flb_vio_params p{};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { p.Rcl[3 * i + j] = Rcl(i, j); p.R_LI[3 * i + j] = Lidar_rot_to_IMU(i, j); }
        p.Pcl[i] = Pcl(i);
        p.t_LI[i] = Lidar_offset_to_IMU(i);
    }
    p.img_point_cov = img_point_cov;
    p.max_iteration = max_iteration;
    p.conv_rot_deg = 0.001f;   // src/lidar_selection.cpp:883
    p.conv_pos_cm = 0.001f;
    p.force_all_passes = 0;
    return p;
}

template <class Image, class SubMap, class StatesGroup>
inline flb_vio_report compute_j(flb_handle* h, const flb_vio_params& prm, const Image& img, SubMap& sub_sparse_map,
                                StatesGroup& state, const StatesGroup& state_propagat) {
    const int total = (int)sub_sparse_map.index.size();
    flb_vio_report rep{};
    if (total == 0) return rep;                                    // :969-970
    std::vector<double> pos;
    std::vector<float> patch;
    std::vector<int> level, src;
    pos.reserve((size_t)total * 3);
    patch.reserve((size_t)total * 192);
    for (int i = 0; i < total; ++i) {
        const auto& pt = sub_sparse_map.voxel_points[i];
        if (pt == nullptr) continue;                               // :798
        for (int k = 0; k < 3; ++k) pos.push_back(pt->pos_(k));
        patch.insert(patch.end(), sub_sparse_map.patch[i].begin(), sub_sparse_map.patch[i].begin() + 192);
        level.push_back(sub_sparse_map.search_levels[i]);
        src.push_back(i);
    }
    check(h, flb_image_upload(h, img.data, img.cols, img.rows, (int)img.step));
    check(h, flb_patches_upload(h, pos.data(), patch.data(), level.data(), (int)level.size()));
    flb_state18 x, xp;
    to_abi(state, x);
    to_abi(state_propagat, xp);
    check(h, flb_vio_update(h, &prm, &x, &xp, &rep));
    from_abi(x, state);
    // sub_sparse_map->errors[i] = patch_error of the last executed pass (:851; display_keypatch reads it, :995)
    std::vector<float> err(level.size());
    check(h, flb_vio_errors(h, err.data(), (int)err.size()));
    for (size_t k = 0; k < src.size(); ++k) sub_sparse_map.errors[src[k]] = err[k];
    return rep;
}

// ---- VIO: one pyramid level, float LidarSelector::UpdateState(cv::Mat img, float total_residual, int level)
// (include/lidar_selection.h:72, src/lidar_selection.cpp:743-902).  The caller keeps its own ComputeJ loop (:974-981):
//     for (level = 2; level >= 0; level--) now_error = flb::update_state(h, prm, img, *sub_sparse_map, *state, *state_propagat, error, level, G);
//     if (now_error < error) state->cov -= G * state->cov;
// G: the caller's 18 x 18 member (anything with operator()(i, j)); its first six columns are rewritten whenever a pass
// of this call was accepted, exactly as :874 does.  The image / patch upload happens on the first level only
// (level == 2) unless `upload` says otherwise.
template <class Image, class SubMap, class StatesGroup, class Mat18>
inline float update_state(flb_handle* h, const flb_vio_params& prm, const Image& img, SubMap& sub_sparse_map, StatesGroup& state,
                          const StatesGroup& state_propagat, float total_residual, int level, Mat18& G, int upload = -1) {
    const int total = (int)sub_sparse_map.index.size();
    if (total == 0) return 0.f;                                    // :746
    std::vector<int> src;
    for (int i = 0; i < total; ++i)
        if (sub_sparse_map.voxel_points[i] != nullptr) src.push_back(i);
    if (upload > 0 || (upload < 0 && level == 2)) {
        std::vector<double> pos;
        std::vector<float> patch;
        std::vector<int> lv;
        for (int i : src) {
            for (int k = 0; k < 3; ++k) pos.push_back(sub_sparse_map.voxel_points[i]->pos_(k));
            patch.insert(patch.end(), sub_sparse_map.patch[i].begin(), sub_sparse_map.patch[i].begin() + 192);
            lv.push_back(sub_sparse_map.search_levels[i]);
        }
        check(h, flb_image_upload(h, img.data, img.cols, img.rows, (int)img.step));
        check(h, flb_patches_upload(h, pos.data(), patch.data(), lv.data(), (int)lv.size()));
    }
    flb_state18 x, xp;
    to_abi(state, x);
    to_abi(state_propagat, xp);
    float last_error = 0.f;
    double Gc[108];
    flb_vio_report rep{};
    check(h, flb_vio_update_level(h, &prm, level, total_residual, &x, &xp, &last_error, Gc, &rep));
    from_abi(x, state);
    if (rep.passes[level] > 0 && last_error < total_residual)     // at least one pass was accepted: G was rewritten (:874)
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 6; ++j) G(i, j) = Gc[6 * i + j];
    std::vector<float> err(src.size());
    check(h, flb_vio_errors(h, err.data(), (int)err.size()));
    for (size_t k = 0; k < src.size(); ++k) sub_sparse_map.errors[src[k]] = err[k];
    return last_error;
}

// ---- IMU forward propagation + backward undistortion (src/IMU_Processing.cpp:655-808) ----------------
// ImuDeque  : the reference's `v_imu` (std::deque<sensor_msgs::Imu::ConstPtr>, last_imu_ pushed to the front,
//             :617-618): elements dereference to .header.stamp.toSec(), .angular_velocity.{x,y,z},
//             .linear_acceleration.{x,y,z}
// PointCloud: pcl_out (:623-639), already cut to this update's points, in time order; compensated in place.
// `carry` holds ImuProcess::last_lidar_end_time_, acc_s_last, angvel_last (the caller keeps `last_imu_ =
// v_imu.back()` itself, :760).  On return state_inout is the propagated state (rot/pos/vel_end, cov).
template <class Mat3, class Vec3>
inline flb_imu_params make_imu_params(const Vec3& cov_gyr, const Vec3& cov_acc, const Vec3& cov_bias_gyr, const Vec3& cov_bias_acc,
                                      double G_m_s2, double mean_acc_norm, const Mat3& Lid_rot_to_IMU, const Vec3& Lid_offset_to_IMU) {
    flb_imu_params p{};
    for (int i = 0; i < 3; ++i) {
        p.cov_gyr[i] = cov_gyr(i); p.cov_acc[i] = cov_acc(i); p.cov_bias_gyr[i] = cov_bias_gyr(i); p.cov_bias_acc[i] = cov_bias_acc(i);
        p.t_LI[i] = Lid_offset_to_IMU(i);
        for (int j = 0; j < 3; ++j) p.R_LI[3 * i + j] = Lid_rot_to_IMU(i, j);
    }
    p.G_m_s2 = G_m_s2;
    p.mean_acc_norm = mean_acc_norm;
    return p;
}

template <class ImuDeque, class StatesGroup, class PointCloud>
inline void undistort_pcl(flb_handle* h, const flb_imu_params& prm, flb_imu_carry& carry, const ImuDeque& v_imu,
                          double pcl_beg_time, double pcl_end_time, StatesGroup& state_inout, PointCloud& pcl_out) {
    std::vector<flb_imu_sample> s;
    s.reserve(v_imu.size());
    for (const auto& m : v_imu) {
        flb_imu_sample e;
        e.t = m->header.stamp.toSec();
        e.gyr[0] = m->angular_velocity.x; e.gyr[1] = m->angular_velocity.y; e.gyr[2] = m->angular_velocity.z;
        e.acc[0] = m->linear_acceleration.x; e.acc[1] = m->linear_acceleration.y; e.acc[2] = m->linear_acceleration.z;
        s.push_back(e);
    }
    flb_state18 x{};
    to_abi(state_inout, x);
    check(h, flb_state_upload(h, &x, &x));
    static_assert(sizeof(pcl_out.points[0]) % sizeof(float) == 0, "point type must be float-packed");
    const int stride = (int)(sizeof(pcl_out.points[0]) / sizeof(float));
    const int n = (int)pcl_out.points.size();
    const int off = n ? (int)(&pcl_out.points[0].curvature - &pcl_out.points[0].x) : 3;
    std::vector<float> out((size_t)n * 3 + 1);
    check(h, flb_imu_undistort(h, &prm, &carry, s.data(), (int)s.size(), pcl_beg_time, pcl_end_time,
                               n ? &pcl_out.points[0].x : nullptr, n ? stride : 4, off, n, out.data(), nullptr, nullptr));
    for (int i = 0; i < n; ++i) {
        pcl_out.points[i].x = out[3 * (size_t)i];
        pcl_out.points[i].y = out[3 * (size_t)i + 1];
        pcl_out.points[i].z = out[3 * (size_t)i + 2];
    }
    check(h, flb_state_download(h, &x, nullptr, nullptr));
    from_abi(x, state_inout);
}

}  // namespace flb
