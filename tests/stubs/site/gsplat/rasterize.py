import torch

from oracle import splat as _os


def rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width, block_width,
                        background=None, return_alpha=False):
    assert block_width == _os.TILE
    if background is None:
        background = torch.ones(colors.shape[-1])
    ids, bins = _os.bin_gaussians(xys, depths, radii, img_height, img_width)
    out, alpha = _os.rasterize(xys, conics, opacity, colors, ids, bins, img_height, img_width, background)
    return (out, alpha) if return_alpha else out
