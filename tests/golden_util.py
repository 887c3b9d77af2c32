import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "T0_golden.npz")


class _Cfg:
    def __init__(self, a):
        self.laser_point_cov, self.img_point_cov, self.cell_size = float(a[0]), float(a[1]), float(a[2])


def load_golden():
    g = dict(np.load(GOLD))
    c = g["cam"]
    frame = {k: g[k] for k in ("map_xyz", "scan_body", "R_prop", "p_prop", "cov", "vel", "bg", "ba", "grav", "R_LI", "t_LI",
                               "Rcl", "Pcl", "image", "patch_pos", "patch_ref", "patch_level")}
    frame["cam"] = dict(width=int(c[0]), height=int(c[1]), fx=c[2], fy=c[3], cx=c[4], cy=c[5], d=tuple(c[6:11]))
    frame["cfg"] = _Cfg(g["cfg"])
    return frame, g
