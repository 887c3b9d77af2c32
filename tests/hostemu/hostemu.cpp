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
