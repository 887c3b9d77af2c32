// Drives flb::undistort_pcl (fastlivo_b200_adapters.hpp) with mock stand-ins for the reference's
// sensor_msgs::Imu / pcl cloud / StatesGroup types.  argv[1]: input blob, argv[2]: output blob
// (formats: tests/test_gpu_adapters.py::test_imu_adapter_matches_oracle).
#include <cstdio>
#include <deque>
#include <memory>
#include <vector>

#include "../../include/fastlivo_b200.h"
#include "../../fast-livo_b200/adapters/fastlivo_b200_adapters.hpp"

template <int R, int C>
struct Mat {   // column-major, like Eigen's default
    double v[R * C] = {0};
    double& operator()(int i, int j) { return v[j * R + i]; }
    double operator()(int i, int j) const { return v[j * R + i]; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};
struct StatesGroup { Mat<3, 3> rot_end; Mat<3, 1> pos_end, vel_end, bias_g, bias_a, gravity; Mat<18, 18> cov; };
struct PointXYZINormal { float x, y, z, pad0, nx, ny, nz, pad1, intensity, curvature, pad2, pad3; };
struct Cloud { std::vector<PointXYZINormal> points; };
struct Stamp { double s; double toSec() const { return s; } };
struct Header { Stamp stamp; };
struct V3 { double x, y, z; };
struct ImuMsg { Header header; V3 angular_velocity, linear_acceleration; };

template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    auto hdr = rd<int>(f, 2);                         // K n
    const int K = hdr[0], n = hdr[1];
    auto imu = rd<double>(f, (size_t)K * 7);
    auto times = rd<double>(f, 3);                    // pcl_beg pcl_end last_lidar_end
    auto st = rd<double>(f, 9 + 15 + 324);            // rot pos vel bg ba grav cov
    auto carryv = rd<double>(f, 6);                   // acc_s_last angvel_last
    auto prm = rd<double>(f, 12 + 2 + 12);            // cov_gyr cov_acc cov_bias_gyr cov_bias_acc, G norm, R_LI t_LI
    auto pts = rd<float>(f, (size_t)n * 4);           // x y z offset_ms
    fclose(f);
    flb_handle* h = nullptr;
    flb_config cfg{};
    if (flb_create(&cfg, &h) != FLB_OK) { fprintf(stderr, "flb_create: %s\n", flb_last_error(nullptr)); return 3; }
    try {
        std::deque<std::shared_ptr<const ImuMsg>> v_imu;
        for (int k = 0; k < K; ++k) {
            auto m = std::make_shared<ImuMsg>();
            m->header.stamp.s = imu[7 * k];
            m->angular_velocity = V3{imu[7 * k + 1], imu[7 * k + 2], imu[7 * k + 3]};
            m->linear_acceleration = V3{imu[7 * k + 4], imu[7 * k + 5], imu[7 * k + 6]};
            v_imu.push_back(m);
        }
        StatesGroup s;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) s.rot_end(i, j) = st[3 * i + j];
            s.pos_end(i) = st[9 + i]; s.vel_end(i) = st[12 + i]; s.bias_g(i) = st[15 + i]; s.bias_a(i) = st[18 + i]; s.gravity(i) = st[21 + i];
        }
        for (int i = 0; i < 18; ++i) for (int j = 0; j < 18; ++j) s.cov(i, j) = st[24 + 18 * i + j];
        Mat<3, 1> cg, ca, cbg, cba, tLI;
        Mat<3, 3> RLI;
        for (int i = 0; i < 3; ++i) {
            cg(i) = prm[i]; ca(i) = prm[3 + i]; cbg(i) = prm[6 + i]; cba(i) = prm[9 + i]; tLI(i) = prm[23 + i];
            for (int j = 0; j < 3; ++j) RLI(i, j) = prm[14 + 3 * i + j];
        }
        flb_imu_params ip = flb::make_imu_params(cg, ca, cbg, cba, prm[12], prm[13], RLI, tLI);
        flb_imu_carry carry{};
        carry.last_lidar_end_time = times[2];
        for (int i = 0; i < 3; ++i) { carry.acc_s_last[i] = carryv[i]; carry.angvel_last[i] = carryv[3 + i]; }
        Cloud pc;
        pc.points.resize(n);
        for (int i = 0; i < n; ++i) {
            pc.points[i] = PointXYZINormal{};
            pc.points[i].x = pts[4 * i]; pc.points[i].y = pts[4 * i + 1]; pc.points[i].z = pts[4 * i + 2]; pc.points[i].curvature = pts[4 * i + 3];
        }
        flb::undistort_pcl(h, ip, carry, v_imu, times[0], times[1], s, pc);
        FILE* o = fopen(argv[2], "wb");
        double out[9 + 6 + 324 + 7];
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) out[3 * i + j] = s.rot_end(i, j); out[9 + i] = s.pos_end(i); out[12 + i] = s.vel_end(i); }
        for (int i = 0; i < 18; ++i) for (int j = 0; j < 18; ++j) out[15 + 18 * i + j] = s.cov(i, j);
        out[339] = carry.last_lidar_end_time;
        for (int i = 0; i < 3; ++i) { out[340 + i] = carry.acc_s_last[i]; out[343 + i] = carry.angvel_last[i]; }
        fwrite(out, 8, 346, o);
        for (int i = 0; i < n; ++i) fwrite(&pc.points[i].x, 4, 3, o);
        fclose(o);
    } catch (const flb::Error& e) {
        fprintf(stderr, "flb error %d: %s\n", e.code, e.what());
        return 4;
    }
    flb_destroy(h);
    return 0;
}
