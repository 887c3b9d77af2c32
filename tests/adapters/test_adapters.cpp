// TEST: the header-only adapters (fast-livo_b200/adapters/fastlivo_b200_adapters.hpp) against mock
// Eigen / PCL / OpenCV-shaped types.  Reads a frame dumped by the pytest driver, runs the
// adapters against libfastlivo_b200.so and writes the results back for comparison with the oracle.
// Mock types mirror only the members the adapters touch (Eigen is column-major: so are the mocks).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../fast-livo_b200/adapters/fastlivo_b200_adapters.hpp"

template <int R, int C>
struct Mat {   // column-major, like Eigen's default
    double v[R * C] = {0};
    double& operator()(int i, int j) { return v[j * R + i]; }
    double operator()(int i, int j) const { return v[j * R + i]; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};
struct StatesGroup {   // include/common_lib.h:296-381
    Mat<3, 3> rot_end;
    Mat<3, 1> pos_end, vel_end, bias_g, bias_a, gravity;
    Mat<18, 18> cov;
};
struct PointXYZINormal { float x, y, z, pad0, nx, ny, nz, pad1, intensity, curvature, pad2, pad3; };
struct Cloud { std::vector<PointXYZINormal> points; };
struct MatX {          // MatrixXd / VectorXd stand-in (column-major)
    int r = 0, c = 1;
    std::vector<double> v;
    void resize(int rr, int cc = 1) { r = rr; c = cc; v.assign((size_t)rr * cc, 0.0); }
    double& operator()(int i, int j) { return v[(size_t)j * r + i]; }
    double& operator()(int i) { return v[i]; }
};
struct DynShare { bool valid = true, converge = true; MatX h_x, h; };
struct Quat {          // MTK SO3 stand-in: only toRotationMatrix()
    Mat<3, 3> R;
    Mat<3, 3> toRotationMatrix() const { return R; }
};
struct StateIkfom { Mat<3, 1> pos; Quat rot, offset_R_L_I; Mat<3, 1> offset_T_L_I; };
struct Point { Mat<3, 1> pos_; };
struct SubSparseMap {
    std::vector<int> index, search_levels;
    std::vector<std::vector<float>> patch;
    std::vector<std::shared_ptr<Point>> voxel_points;
    std::vector<float> errors;
};
struct Image { unsigned char* data; int cols, rows; size_t step; };

template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    auto hdr = rd<int>(f, 6);   // M N Pn W H T
    const int M = hdr[0], N = hdr[1], Pn = hdr[2], W = hdr[3], H = hdr[4], T = hdr[5];
    auto map = rd<float>(f, (size_t)M * 3);
    auto scan = rd<float>(f, (size_t)N * 3);
    auto st = rd<double>(f, 9 + 3 + 324 + 3);            // R_prop p_prop cov grav
    auto ext = rd<double>(f, 9 + 3 + 9 + 3);             // R_LI t_LI Rcl Pcl
    auto camv = rd<double>(f, 11);
    auto covs = rd<double>(f, 2);                        // laser_point_cov img_point_cov
    auto img = rd<unsigned char>(f, (size_t)W * H);
    auto ppos = rd<double>(f, (size_t)Pn * 3);
    auto pref = rd<float>(f, (size_t)Pn * 192);
    auto plev = rd<int>(f, Pn);
    fclose(f);

    flb_handle* h = nullptr;
    flb_config cfg{};
    cfg.persistent = 1;
    if (flb_create(&cfg, &h) != FLB_OK) { fprintf(stderr, "flb_create: %s\n", flb_last_error(nullptr)); return 3; }
    try {
        std::vector<PointXYZINormal> mp(M);
        for (int i = 0; i < M; ++i) { mp[i] = PointXYZINormal{}; mp[i].x = map[3 * i]; mp[i].y = map[3 * i + 1]; mp[i].z = map[3 * i + 2]; }
        flb::upload_map(h, mp);
        Cloud body;
        body.points.resize(N);
        for (int i = 0; i < N; ++i) { body.points[i] = PointXYZINormal{}; body.points[i].x = scan[3 * i]; body.points[i].y = scan[3 * i + 1]; body.points[i].z = scan[3 * i + 2]; }
        flb::upload_scan(h, body);

        StatesGroup state;
        Mat<3, 3> R_LI, Rcl;
        Mat<3, 1> t_LI, Pcl;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) { state.rot_end(i, j) = st[3 * i + j]; R_LI(i, j) = ext[3 * i + j]; Rcl(i, j) = ext[12 + 3 * i + j]; }
            state.pos_end(i) = st[9 + i];
            state.gravity(i) = st[9 + 3 + 324 + i];
            t_LI(i) = ext[9 + i];
            Pcl(i) = ext[21 + i];
        }
        for (int i = 0; i < 18; ++i) for (int j = 0; j < 18; ++j) state.cov(i, j) = st[12 + 18 * i + j];
        StatesGroup prop = state;

        // (1) IKFoM-typed callback: rows of one pass
        StateIkfom s;
        s.pos = state.pos_end; s.rot.R = state.rot_end; s.offset_R_L_I.R = R_LI; s.offset_T_L_I = t_LI;
        DynShare ds;
        int effct = 0;
        double res_mean = 0;
        std::vector<int> eidx;
        flb::h_share_model(h, covs[0], s, ds, effct, res_mean, &eidx);

        // (2) live LIO update
        flb_lio_params lp = flb::make_lio_params(R_LI, t_LI, covs[0], T);
        flb_lio_report lrep = flb::lio_update(h, lp, state, prop);

        // (3) ComputeJ
        flb_camera cam{};
        cam.width = (int)camv[0]; cam.height = (int)camv[1]; cam.fx = camv[2]; cam.fy = camv[3]; cam.cx = camv[4]; cam.cy = camv[5];
        for (int i = 0; i < 5; ++i) cam.d[i] = camv[6 + i];
        flb::check(h, flb_camera_set(h, &cam));
        SubSparseMap sub;
        for (int i = 0; i < Pn; ++i) {
            sub.index.push_back(i);
            sub.search_levels.push_back(plev[i]);
            sub.patch.emplace_back(pref.begin() + (size_t)i * 192, pref.begin() + (size_t)(i + 1) * 192);
            auto p = std::make_shared<Point>();
            for (int k = 0; k < 3; ++k) p->pos_(k) = ppos[3 * i + k];
            sub.voxel_points.push_back(i % 17 == 5 ? nullptr : p);   // a few null points, like the reference's map
            sub.errors.push_back(-1.0f);                              // :569 (overwritten by every pass, :851)
        }
        Image im{img.data(), W, H, (size_t)W};
        flb_vio_params vp = flb::make_vio_params(Rcl, Pcl, R_LI, t_LI, covs[1], T);
        StatesGroup vprop = state;
        // the same ComputeJ assembled from three UpdateState calls, as the reference's own loop does (:974-981)
        StatesGroup state2 = state;
        SubSparseMap sub2 = sub;
        Mat<18, 18> G;
        float now_error = 1e10f;
        for (int level = 2; level >= 0; level--) now_error = flb::update_state(h, vp, im, sub2, state2, vprop, 1e10f, level, G);
        if (now_error < 1e10f) {
            Mat<18, 18> GP;
            for (int i = 0; i < 18; ++i)
                for (int j = 0; j < 18; ++j) {
                    double acc = 0;
                    for (int k = 0; k < 18; ++k) acc += G(i, k) * state2.cov(k, j);
                    GP(i, j) = acc;
                }
            for (int i = 0; i < 18; ++i)
                for (int j = 0; j < 18; ++j) state2.cov(i, j) -= GP(i, j);
        }
        flb_vio_report vrep = flb::compute_j(h, vp, im, sub, state, vprop);
        double lvl_diff = 0;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) lvl_diff = std::max(lvl_diff, std::fabs(state.rot_end(i, j) - state2.rot_end(i, j)));
            lvl_diff = std::max(lvl_diff, std::fabs(state.pos_end(i) - state2.pos_end(i)));
        }
        double cov_diff = 0, cov_max = 0;
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) {
                cov_diff = std::max(cov_diff, std::fabs(state.cov(i, j) - state2.cov(i, j)));
                cov_max = std::max(cov_max, std::fabs(state.cov(i, j)));
            }
        bool err_same = true;
        for (int i = 0; i < Pn; ++i) err_same = err_same && (sub.errors[i] == sub2.errors[i]);
        if (lvl_diff > 1e-12 || cov_diff > 1e-9 * cov_max || !err_same) {
            fprintf(stderr, "update_state x3 != compute_j: state %.3e cov %.3e errors %d\n", lvl_diff, cov_diff, (int)err_same);
            return 5;
        }

        FILE* o = fopen(argv[2], "wb");
        int oh[8] = {effct, lrep.passes, lrep.n_eff_last, vrep.passes[0], vrep.passes[1], vrep.passes[2], vrep.cov_updated, (int)eidx.size()};
        fwrite(oh, sizeof(int), 8, o);
        fwrite(&res_mean, sizeof(double), 1, o);
        // h_x back to row-major for the checker
        for (int r = 0; r < effct; ++r) for (int c = 0; c < 12; ++c) { double v = ds.h_x(r, c); fwrite(&v, 8, 1, o); }
        fwrite(ds.h.v.data(), 8, effct, o);
        fwrite(eidx.data(), sizeof(int), eidx.size(), o);
        double out[12 + 324];
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) out[3 * i + j] = state.rot_end(i, j); out[9 + i] = state.pos_end(i); }
        for (int i = 0; i < 18; ++i) for (int j = 0; j < 18; ++j) out[12 + 18 * i + j] = state.cov(i, j);
        fwrite(out, 8, 12 + 324, o);
        fwrite(sub.errors.data(), sizeof(float), sub.errors.size(), o);
        fclose(o);
    } catch (const flb::Error& e) {
        fprintf(stderr, "flb error %d: %s\n", e.code, e.what());
        return 4;
    }
    flb_destroy(h);
    return 0;
}
