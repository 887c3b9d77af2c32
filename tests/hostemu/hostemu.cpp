// TEST INFRASTRUCTURE -- compiles the product's per-thread device math
// (fast-livo_b200/csrc/flb_device.cuh) with g++ so that the CPU-only test tier can
// check it against the oracle without a GPU.  This library is NEVER loaded by the
// product; it exists only under tests/.
#include "../../fast-livo_b200/csrc/flb_device.cuh"
#include <cstring>

using namespace flb;

extern "C" {

int emu_plane_fit(const float* nb15, float thr, float* out4) {
    float nb[kMatch][3];
    for (int j = 0; j < kMatch; ++j)
        for (int c = 0; c < 3; ++c) nb[j][c] = nb15[3 * j + c];
    float pabcd[4];
    const bool ok = plane_fit5(nb, thr, pabcd);
    std::memcpy(out4, pabcd, sizeof(pabcd));
    return ok ? 1 : 0;
}

static void make_pose(const double* R, const double* p, const double* R_LI, const double* t_LI, LioPose& pose) {
    for (int i = 0; i < 9; ++i) { pose.R[i] = R[i]; pose.R_LI[i] = R_LI[i]; }
    for (int i = 0; i < 3; ++i) { pose.p[i] = p[i]; pose.t_LI[i] = t_LI[i]; }
    m3_T(pose.R, pose.Rt);
    m3_T(pose.R_LI, pose.RLIt);
}

// per point: world, pd2, gate, rows (given the plane)
void emu_lio_points(const double* R, const double* p, const double* R_LI, const double* t_LI, const float* body, int N,
                    const float* pabcd, float* world, float* pd2, unsigned char* gate, double* row6, double* row12) {
    LioPose pose;
    make_pose(R, p, R_LI, t_LI, pose);
    for (int i = 0; i < N; ++i) {
        double pI[3];
        float pw[3];
        lio_body_to_world(pose, body + 3 * i, pI, pw);
        std::memcpy(world + 3 * i, pw, sizeof(pw));
        float d;
        gate[i] = lio_residual(pabcd + 4 * i, pw, body + 3 * i, &d) ? 1 : 0;
        pd2[i] = d;
        lio_row6(pose, pI, pabcd + 4 * i, row6 + 6 * i);
        lio_row12(pose, pI, body + 3 * i, pabcd + 4 * i, row12 + 12 * i);
    }
}

void emu_knn(const float* gridf /*ox,oy,oz,cell,inv_cell,max_d2*/, const int* gridi /*nx,ny,nz,max_ring*/,
             const int* cell_start, const float* pts4, const float* q, int nq, int* pos, float* d2) {
    GridDesc g;
    g.ox = gridf[0]; g.oy = gridf[1]; g.oz = gridf[2]; g.cell = gridf[3]; g.inv_cell = gridf[4]; g.max_d2 = gridf[5];
    g.nx = gridi[0]; g.ny = gridi[1]; g.nz = gridi[2]; g.max_ring = gridi[3];
    const map_pt* pts = reinterpret_cast<const map_pt*>(pts4);
    for (int i = 0; i < nq; ++i) {
        Top5 t;
        int rows[18];
        knn5_grid(g, cell_start, pts, q[3 * i], q[3 * i + 1], q[3 * i + 2], t, rows, 1);
        for (int j = 0; j < kMatch; ++j) { pos[5 * i + j] = t.i[j]; d2[5 * i + j] = t.d[j]; }
    }
}

struct LatHost {
    const float* p;
    float operator()(int r, int c) const { return p[r * 11 + c]; }
};

// cam: width,height as doubles then fx,fy,cx,cy,d0..d4 ; returns per-patch validity
void emu_vio(const double* camv, const double* Rci, const double* Pci, const double* R, const double* p, const double* pos,
             const float* patch, const int* search_level, int Pn, int level, const unsigned char* img, double* z, double* H,
             float* errors, unsigned char* valid) {
    CamModel cam;
    cam.width = (int)camv[0]; cam.height = (int)camv[1];
    cam.fx = camv[2]; cam.fy = camv[3]; cam.cx = camv[4]; cam.cy = camv[5];
    for (int i = 0; i < 5; ++i) cam.d[i] = camv[6 + i];
    cam.jfx = fabs(cam.fx);
    cam.jfy = fabs(4.0 * cam.fx * cam.fy) / (4. * cam.jfx);
    VioPose vp;
    double Jdphi_dR[9], Jdp_dR[9];
    vio_pose_constants(Rci, Pci, Jdphi_dR, Jdp_dR);
    vio_make_pose(Rci, Pci, Jdphi_dR, Jdp_dR, R, p, vp);
    for (int i = 0; i < Pn; ++i) {
        PatchGeom g;
        vio_patch_geom(cam, vp, pos + 3 * i, level, search_level[i], g);
        valid[i] = g.valid ? 1 : 0;
        errors[i] = 0.f;
        for (int e = 0; e < 64; ++e) { z[64 * i + e] = 0.0; for (int r = 0; r < 6; ++r) H[(64 * i + e) * 6 + r] = 0.0; }
        if (!g.valid) continue;
        float lat[121];
        const int W = cam.width;
        const unsigned char* base = img + (size_t)(g.v_i - 5 * g.scale) * W + (g.u_i - 5 * g.scale);
        for (int e = 0; e < 121; ++e) {
            const int r = e / 11, c = e - r * 11;
            lat[e] = (float)base[(size_t)r * g.scale * W + c * g.scale];
        }
        LatHost L{lat};
        float pe = 0.f;
        for (int x = 0; x < 8; ++x)
            for (int y = 0; y < 8; ++y) {
                double row[6], res;
                vio_pixel(L, g, vp, x, y, patch[192 * i + 64 * level + x * 8 + y], row, &res);
                z[64 * i + x * 8 + y] = res;
                std::memcpy(H + (64 * i + x * 8 + y) * 6, row, sizeof(row));
                pe = (float)((double)pe + res * res);
            }
        errors[i] = pe;
    }
}

// row f4: per-point candidate of addSparseMap's first loop; cell[i] = -1 when the point is not in the frame
void emu_visual_candidates(const double* camv, const double* Rcw, const double* Pcw, const unsigned char* img, const float* xyz,
                           int n, int grid_size, int border, int* cell, float* score) {
    CamModel cam;
    cam.width = (int)camv[0]; cam.height = (int)camv[1];
    cam.fx = camv[2]; cam.fy = camv[3]; cam.cx = camv[4]; cam.cy = camv[5];
    for (int i = 0; i < 5; ++i) cam.d[i] = camv[6 + i];
    cam.jfx = cam.jfy = 0.0;
    const int gh = cam.height / grid_size;
    for (int i = 0; i < n; ++i) {
        score[i] = 0.0f;
        cell[i] = visual_candidate(cam, Rcw, Pcw, img, cam.width, xyz + 3 * i, grid_size, gh, border, &score[i]);
    }
}
float emu_shi_tomasi(const unsigned char* img, int w, int h, int u, int v) { return shi_tomasi_score(img, w, h, w, u, v); }

// row f3: the per-point backward compensation with pose `head` (22 doubles: t, acc, gyr, vel, pos, rot)
void emu_imu_compensate(const double* head22, const double* R_LI, const double* t_LI, const double* rot_end, const double* pos_end,
                        const float* xyz, const float* offset_ms, int n, float* out) {
    ImuPose hp;
    std::memcpy(&hp, head22, sizeof(hp));
    double RLIt[9], RendT[9], extR_Ri[9], exrR_extT[3];
    m3_T(R_LI, RLIt);
    m3_T(rot_end, RendT);
    m3_mul(RLIt, RendT, extR_Ri);
    m3_vec(RLIt, t_LI, exrR_extT);
    for (int i = 0; i < n; ++i) {
        float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        imu_compensate_point(hp, R_LI, t_LI, extR_Ri, exrR_extT, pos_end, (double)offset_ms[i] / double(1000), p);
        out[3 * i] = p[0]; out[3 * i + 1] = p[1]; out[3 * i + 2] = p[2];
    }
}
void emu_exp_w_dt(const double* w, double dt, double* R) { so3_exp_dt(w, dt, R); }

void emu_exp3(const double* v, double* R) { so3_exp(v, R); }
void emu_log3(const double* R, double* o) { so3_log(R, o); }

}  // extern "C"

// ---- IKFoM manifold algebra (row a8) ----------------------------------------------------------
extern "C" {
void emu_ikfom_boxplus(double* state26, const double* d23) {
    StateIkfom x;
    std::memcpy(&x, state26, 26 * sizeof(double));
    ikfom_boxplus(x, d23);
    std::memcpy(state26, &x, 26 * sizeof(double));
}
void emu_ikfom_boxminus(const double* a26, const double* b26, double* res23) {
    StateIkfom a, b;
    std::memcpy(&a, a26, 26 * sizeof(double));
    std::memcpy(&b, b26, 26 * sizeof(double));
    ikfom_boxminus(a, b, res23);
}
void emu_ikfom_proj(const double* grav_x, const double* grav_prop, const double* seg2, const double* seg3, double* T2, double* AT) {
    double Nx[6], Mx[6], A[9];
    s2_Nx_yy(grav_x, Nx);
    s2_Mx(grav_prop, seg2, Mx);
    mm_small(Nx, 2, 3, Mx, 2, T2);
    mtk_A_matrix(seg3, A);
    transpose_small(A, 3, 3, AT);
}
}

// ---- visual map: selection + warp (row f2), growth + observations (row f4) ---------------------------------
// A sequential stand-in for the device kernels of fast-livo_b200/csrc (same per-element functions, the atomics
// replaced by their sequential meaning), so that the CPU tier can check the device math against the oracle.
#include <unordered_set>
#include <vector>

namespace {
struct HostCtx {
    int lane = 0, n = 1;
    void sync() const {}
    bool any(bool b) const { return b; }
};
struct EmuVm {
    CamModel cam;
    VmParams prm;
    std::vector<VmPoint> pts;
    std::vector<VmFeature> fts;
    std::vector<unsigned char> pool;
    std::vector<float> map_value;
    int img_frame = -1, img_slot = -1, n_img = 0;
    std::vector<int> sel_index, sel_point, sel_level;
    std::vector<float> sel_error, sel_patch;
};
int emu_store_image(EmuVm* vm, const unsigned char* img, int frame_id) {
    if (vm->img_frame == frame_id && vm->img_slot >= 0) return vm->img_slot;
    const size_t sz = (size_t)vm->cam.width * vm->cam.height;
    vm->pool.insert(vm->pool.end(), img, img + sz);
    vm->img_frame = frame_id;
    vm->img_slot = vm->n_img++;
    return vm->img_slot;
}
void emu_push_feature(EmuVm* vm, int point, const double* pc, const double* Rcw, const double* Pcw, float score, int level, int id, int img) {
    VmFeature ft;
    ft.px[0] = pc[0]; ft.px[1] = pc[1];
    cam2world(vm->cam, pc[0], pc[1], ft.f);
    std::memcpy(ft.R, Rcw, sizeof(ft.R));
    std::memcpy(ft.t, Pcw, sizeof(ft.t));
    ft.score = score; ft.level = level; ft.id = id; ft.img = img;
    vm->fts.push_back(ft);
    VmPoint& p = vm->pts[point];
    for (int q = p.n_obs; q > 0; --q) p.obs[q] = p.obs[q - 1];      // push_front
    p.obs[0] = (int)vm->fts.size() - 1;
    p.n_obs++;
}
}  // namespace

extern "C" {

void* emu_vm_create(const double* camv, int grid_size, double outlier_threshold, int ncc_en, double ncc_thre) {
    EmuVm* vm = new EmuVm;
    CamModel& cam = vm->cam;
    cam.width = (int)camv[0]; cam.height = (int)camv[1];
    cam.fx = camv[2]; cam.fy = camv[3]; cam.cx = camv[4]; cam.cy = camv[5];
    for (int i = 0; i < 5; ++i) cam.d[i] = camv[6 + i];
    cam.jfx = fabs(cam.fx);
    cam.jfy = fabs(4.0 * cam.fx * cam.fy) / (4. * cam.jfx);
    vm->prm.grid_size = grid_size;
    vm->prm.grid_n_width = cam.width / grid_size;
    vm->prm.grid_n_height = cam.height / grid_size;
    vm->prm.length = vm->prm.grid_n_width * vm->prm.grid_n_height;
    vm->prm.halfpatch = 4;
    vm->prm.ncc_en = ncc_en;
    vm->prm.outlier_threshold = outlier_threshold;
    vm->prm.ncc_thre = ncc_thre;
    vm->map_value.assign(vm->prm.length, 0.0f);
    return vm;
}
void emu_vm_destroy(void* h) { delete static_cast<EmuVm*>(h); }
void emu_vm_counts(void* h, int* out3) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    out3[0] = (int)vm->pts.size(); out3[1] = (int)vm->fts.size(); out3[2] = vm->n_img;
}
void emu_vm_map_value(void* h, float* out) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    std::memcpy(out, vm->map_value.data(), sizeof(float) * vm->prm.length);
}

int emu_vm_select(void* h, const unsigned char* img, const double* Rcw, const double* Pcw, const float* pg_down, int n) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    vm->sel_index.clear(); vm->sel_point.clear(); vm->sel_level.clear(); vm->sel_error.clear(); vm->sel_patch.clear();
    if (vm->pts.empty()) return 0;
    const CamModel& cam = vm->cam;
    const VmParams& prm = vm->prm;
    const int width = cam.width, border = (prm.halfpatch + 1) * 8;
    std::fill(vm->map_value.begin(), vm->map_value.end(), 0.0f);
    std::vector<unsigned long long> depth((size_t)width * cam.height, 0ull);
    std::unordered_set<unsigned long long> keys;
    // kernel 1: per scan point -- voxel set + depth image (atomicMax on {index + 1 : depth bits})
    for (int i = 0; i < n; ++i) {
        keys.insert(vm_scan_key(pg_down + 3 * i));
        const double pt_w[3] = {pg_down[3 * i], pg_down[3 * i + 1], pg_down[3 * i + 2]};
        double pt_c[3];
        se3_apply(Rcw, Pcw, pt_w, pt_c);
        if (pt_c[2] > 0) {
            const double px0 = cam.jfx * pt_c[0] / pt_c[2] + cam.cx, px1 = cam.jfy * pt_c[1] / pt_c[2] + cam.cy;
            if (vm_in_frame(cam, px0, px1, border)) {
                const float d = pt_c[2];
                const unsigned long long v = ((unsigned long long)(i + 1) << 32) | float_as_u32(d);
                unsigned long long& c = depth[(size_t)width * (int)px1 + (int)px0];
                if (v > c) c = v;
            }
        }
    }
    // kernel 2: per visual-map point -- cell claims (atomicMin on {dist bits : ~index}, atomicMax on the value)
    double frame_pos[3];
    se3_pos(Rcw, Pcw, frame_pos);
    std::vector<unsigned long long> best(prm.length, ~0ull);
    std::vector<unsigned char> type(prm.length, 0);
    for (size_t j = 0; j < vm->pts.size(); ++j) {
        const VmPoint& pt = vm->pts[j];
        if (!keys.count(vm_pack_key(pt.key[0], pt.key[1], pt.key[2]))) continue;
        double pt_cam[3], pc[2];
        se3_apply(Rcw, Pcw, pt.pos, pt_cam);
        if (pt_cam[2] < 0) continue;
        world2cam(cam, pt_cam, pc);
        if (!vm_in_frame(cam, pc[0], pc[1], border)) continue;
        const int index = (int)(pc[0] / prm.grid_size) * prm.grid_n_height + (int)(pc[1] / prm.grid_size);
        type[index] = 1;
        const double ov[3] = {frame_pos[0] - pt.pos[0], frame_pos[1] - pt.pos[1], frame_pos[2] - pt.pos[2]};
        const float cur_dist = norm3(ov);
        if (cur_dist <= 10000.0f) {
            const unsigned long long v = ((unsigned long long)float_as_u32(cur_dist) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)j);
            if (v < best[index]) best[index] = v;
        }
        if (pt.value >= vm->map_value[index]) vm->map_value[index] = pt.value;
    }
    // kernel 3 + compaction: per cell
    HostCtx cx;
    std::vector<float> patch_wrap(192), patch_cur(64);
    double shared8[8];
    for (int i = 0; i < prm.length; ++i) {
        if (!type[i] || best[i] == ~0ull) continue;
        const int pj = (int)(0xFFFFFFFFu - (unsigned)(best[i] & 0xFFFFFFFFull));
        int level;
        float error;
        if (!vm_build_cell(cx, cam, prm, Rcw, Pcw, frame_pos, vm->pts[pj], vm->fts.data(), vm->pool.data(), img, depth.data(),
                           patch_wrap.data(), patch_cur.data(), shared8, &level, &error))
            continue;
        vm->sel_index.push_back(i);
        vm->sel_point.push_back(pj);
        vm->sel_level.push_back(level);
        vm->sel_error.push_back(error);
        vm->sel_patch.insert(vm->sel_patch.end(), patch_wrap.begin(), patch_wrap.end());
    }
    return (int)vm->sel_index.size();
}

void emu_vm_selected(void* h, int* index, int* point, int* level, float* error, double* pos, float* patch) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    const size_t n = vm->sel_index.size();
    std::memcpy(index, vm->sel_index.data(), n * sizeof(int));
    std::memcpy(point, vm->sel_point.data(), n * sizeof(int));
    std::memcpy(level, vm->sel_level.data(), n * sizeof(int));
    std::memcpy(error, vm->sel_error.data(), n * sizeof(float));
    std::memcpy(patch, vm->sel_patch.data(), n * 192 * sizeof(float));
    for (size_t i = 0; i < n; ++i) std::memcpy(pos + 3 * i, vm->pts[vm->sel_point[i]].pos, 3 * sizeof(double));
}

int emu_vm_grow(void* h, const unsigned char* img, const double* Rcw, const double* Pcw, const float* pg, int n, int frame_id) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    const CamModel& cam = vm->cam;
    const VmParams& prm = vm->prm;
    const int border = (prm.halfpatch + 1) * 8;
    // candidate scoring: {ordered score bits : ~index} atomicMax per cell, seeded with the incoming value
    std::vector<int> winner(prm.length, -1);
    for (int i = 0; i < n; ++i) {
        float score = 0.0f;
        const int cell = visual_candidate(cam, Rcw, Pcw, img, cam.width, pg + 3 * i, prm.grid_size, prm.grid_n_height, border, &score);
        if (cell < 0) continue;
        if (score > vm->map_value[cell]) { vm->map_value[cell] = score; winner[cell] = i; }
    }
    int add = 0;
    for (int c = 0; c < prm.length; ++c) {
        if (winner[c] < 0) continue;
        const double pt[3] = {pg[3 * winner[c]], pg[3 * winner[c] + 1], pg[3 * winner[c] + 2]};
        double pf[3], pc[2];
        se3_apply(Rcw, Pcw, pt, pf);
        world2cam(cam, pf, pc);
        VmPoint p;
        std::memset(&p, 0, sizeof(p));
        std::memcpy(p.pos, pt, sizeof(pt));
        p.value = vm->map_value[c];
        p.n_obs = 0;
        vm_feat_key(pt, p.key);
        vm->pts.push_back(p);
        const int slot = emu_store_image(vm, img, frame_id);
        emu_push_feature(vm, (int)vm->pts.size() - 1, pc, Rcw, Pcw, vm->map_value[c], 0, frame_id, slot);
        add++;
    }
    return add;
}

int emu_vm_add_observations(void* h, const unsigned char* img, const double* Rcw, const double* Pcw, int frame_id) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    double cur_pos[3];
    se3_pos(Rcw, Pcw, cur_pos);
    int added = 0;
    for (size_t i = 0; i < vm->sel_index.size(); ++i) {
        VmPoint& pt = vm->pts[vm->sel_point[i]];
        double pc[2];
        int erase_slot;
        const bool add_flag = vm_observation_test(vm->cam, Rcw, Pcw, cur_pos, pt, vm->fts.data(), pc, &erase_slot);
        if (erase_slot >= 0) {
            for (int q = erase_slot; q + 1 < pt.n_obs; ++q) pt.obs[q] = pt.obs[q + 1];
            pt.n_obs--;
        }
        if (add_flag) {
            pt.value = shi_tomasi_score(img, vm->cam.width, vm->cam.height, vm->cam.width, (int)pc[0], (int)pc[1]);
            const int slot = emu_store_image(vm, img, frame_id);
            emu_push_feature(vm, vm->sel_point[i], pc, Rcw, Pcw, pt.value, vm->sel_level[i], frame_id, slot);
            added++;
        }
    }
    return added;
}

void emu_vm_dump_points(void* h, double* pos, float* value, int* n_obs, int* obs20) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    for (size_t j = 0; j < vm->pts.size(); ++j) {
        std::memcpy(pos + 3 * j, vm->pts[j].pos, 3 * sizeof(double));
        value[j] = vm->pts[j].value;
        n_obs[j] = vm->pts[j].n_obs;
        for (int q = 0; q < 20; ++q) obs20[20 * j + q] = q < vm->pts[j].n_obs ? vm->pts[j].obs[q] : -1;
    }
}
void emu_vm_dump_features(void* h, double* geo17, float* score, int* level_id_img) {
    EmuVm* vm = static_cast<EmuVm*>(h);
    for (size_t j = 0; j < vm->fts.size(); ++j) {
        const VmFeature& f = vm->fts[j];
        double* g = geo17 + 17 * j;
        g[0] = f.px[0]; g[1] = f.px[1];
        std::memcpy(g + 2, f.f, 3 * sizeof(double));
        std::memcpy(g + 5, f.R, 9 * sizeof(double));
        std::memcpy(g + 14, f.t, 3 * sizeof(double));
        score[j] = f.score;
        level_id_img[3 * j] = f.level; level_id_img[3 * j + 1] = f.id; level_id_img[3 * j + 2] = f.img;
    }
}

}  // extern "C"

extern "C" void emu_colorize(const double* camv, const double* Rcw, const double* Pcw, const unsigned char* bgr, const float* xyz, int n,
                             unsigned char* rgb, unsigned char* valid) {
    CamModel cam;
    cam.width = (int)camv[0]; cam.height = (int)camv[1];
    cam.fx = camv[2]; cam.fy = camv[3]; cam.cx = camv[4]; cam.cy = camv[5];
    for (int i = 0; i < 5; ++i) cam.d[i] = camv[6 + i];
    cam.jfx = cam.jfy = 0.0;
    for (int i = 0; i < n; ++i) {
        unsigned char c[3] = {0, 0, 0};
        valid[i] = vm_colorize_point(cam, Rcw, Pcw, bgr, xyz + 3 * i, c) ? 1 : 0;
        rgb[3 * i] = c[0]; rgb[3 * i + 1] = c[1]; rgb[3 * i + 2] = c[2];
    }
}
