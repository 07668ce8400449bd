from oracle import splat as _os


def project_gaussians(means3d, scales, glob_scale, quats, viewmat, fx, fy, cx, cy, img_height, img_width, block_width,
                      clip_thresh=0.01):
    assert block_width == _os.TILE
    p = _os.project_gaussians(means3d, scales, glob_scale, quats, viewmat, fx, fy, cx, cy, img_height, img_width,
                              clip_thresh)
    return p["xys"], p["depths"], p["radii"], p["conics"], p["compensation"], p["num_tiles_hit"], p["cov3d"]
