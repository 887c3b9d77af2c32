/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see flo_oracle.h).
 *
 * flo_visual_candidates: CPU restatement of the scoring loop of LidarSelector::addSparseMap
 * (reference src/lidar_selection.cpp:150-168): every scan point is projected into the
 * new frame, scored with vikit's Shi-Tomasi corner score and competes for its image-grid cell.
 * vikit (uzh-rpg/rpg_vikit, vikit_common/src/vision.cpp: shiTomasiScore; abstract_camera.h: isInFrame;
 * pinhole_camera.cpp: world2cam) is a third-party dependency absent from /root/reference; its published
 * algorithm is restated here.  PARITY STATUS: "parity unpinned".
 */
#include <cmath>
#include <cstdint>

#include "flo_oracle.h"

/* vk::shiTomasiScore(const cv::Mat& img, int u, int v): 8x8 box of central differences, smaller eigenvalue of
 * the (normalised) structure tensor.  All sums are integers < 2^24, i.e. exact in float. */
extern "C" float flo_shi_tomasi(const uint8_t* img, int width, int height, int stride, int u, int v) {
    float dXX = 0.0f, dYY = 0.0f, dXY = 0.0f;
    const int halfbox_size = 4;
    const int box_size = 2 * halfbox_size;
    const int box_area = box_size * box_size;
    const int x_min = u - halfbox_size, x_max = u + halfbox_size, y_min = v - halfbox_size, y_max = v + halfbox_size;
    if (x_min < 1 || x_max >= width - 1 || y_min < 1 || y_max >= height - 1) return 0.0f;   /* too close to the border */
    for (int y = y_min; y < y_max; ++y) {
        const uint8_t* ptr_left = img + stride * y + x_min - 1;
        const uint8_t* ptr_right = img + stride * y + x_min + 1;
        const uint8_t* ptr_top = img + stride * (y - 1) + x_min;
        const uint8_t* ptr_bottom = img + stride * (y + 1) + x_min;
        for (int x = 0; x < box_size; ++x, ++ptr_left, ++ptr_right, ++ptr_top, ++ptr_bottom) {
            const float dx = (float)(*ptr_right - *ptr_left);
            const float dy = (float)(*ptr_bottom - *ptr_top);
            dXX += dx * dx;
            dYY += dy * dy;
            dXY += dx * dy;
        }
    }
    /* find and return the smaller eigenvalue */
    dXX = (float)(dXX / (2.0 * box_area));
    dYY = (float)(dYY / (2.0 * box_area));
    dXY = (float)(dXY / (2.0 * box_area));
    return (float)(0.5 * (dXX + dYY - std::sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY))));
}

/* addSparseMap's first loop, src/lidar_selection.cpp:150-168.  map_value (length = grid_n_width * grid_n_height)
 * is in/out: addFromSparseMap has already written the scores of the visible map points into it (:447-455).
 * winner[cell] = index of the scan point that took the cell in this call, -1 otherwise. */
extern "C" void flo_visual_candidates(const flo_cam* cam, const double* Rcw, const double* Pcw, const uint8_t* img, int stride,
                                      const float* world_xyz, int n, int grid_size, int border, float* map_value,
                                      int* winner) {
    const int grid_n_width = cam->width / grid_size, grid_n_height = cam->height / grid_size;   /* :55-56 */
    for (int c = 0; c < grid_n_width * grid_n_height; ++c) winner[c] = -1;
    for (int i = 0; i < n; ++i) {
        const double pt[3] = {world_xyz[3 * i], world_xyz[3 * i + 1], world_xyz[3 * i + 2]};
        double pf[3], pc[2];
        for (int k = 0; k < 3; ++k) pf[k] = Rcw[3 * k] * pt[0] + Rcw[3 * k + 1] * pt[1] + Rcw[3 * k + 2] * pt[2] + Pcw[k];   /* T_f_w * pt */
        flo_world2cam(cam, pf, pc);                                                  /* Frame::w2c, include/frame.h:89 */
        const int ox = (int)pc[0], oy = (int)pc[1];                                 /* pc.cast<int>() */
        if (!(std::fabs(pc[0]) < 1e9) || !(std::fabs(pc[1]) < 1e9)) continue;      /* cast of a non-finite / huge value is UB in the reference */
        if (!(ox >= border && ox < cam->width - border && oy >= border && oy < cam->height - border)) continue;   /* isInFrame, :154 */
        const int index = (int)(pc[0] / grid_size) * grid_n_height + (int)(pc[1] / grid_size);   /* :156 */
        const float cur_value = flo_shi_tomasi(img, cam->width, cam->height, stride, (int)pc[0], (int)pc[1]);   /* :158 */
        if (cur_value > map_value[index]) {                                          /* :160 */
            map_value[index] = cur_value;
            winner[index] = i;
        }
    }
}
