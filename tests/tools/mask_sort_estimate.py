"""How many kernel offsets does a 128-row tile of the implicit GEMM touch?  (CPU estimate, uses the oracle.)

The tensor-core sparse conv (csrc/spconv_tc.cu) skips a K chunk only when NO row of the tile uses its
kernel offsets.  With rows in index-set order (first appearance at level 0, ascending linear index below)
nearly every tile touches nearly all 27 offsets although a row uses 12-50 % of them.  spconv-2.x sorts the
output rows by their 27-bit neighbour mask before tiling (`mask_argsort_fwd_splits`, call site
bug_fix/conv.py:382-415).  This script measures what that ordering would buy on the bench scenes: average
active offsets per tile, current order vs rows sorted by mask, for the four SubM resolution levels of the
LiDAR encoder, for a sort by the full mask (spconv's choice) and by a 15-bit structural digest
(two 8-bit radix passes).  Output committed as profiles/r01g_mask_sort_estimate.txt.

    python tests/tools/mask_sort_estimate.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from msmdfusion_b200 import synthetic  # noqa: E402
from oracle import cpu  # noqa: E402  (analysis script kept under tests/: it uses the oracle as its rulebook source)


def tiles_active(used, order):
    n = used.shape[1]
    nt = (n + 127) // 128
    u = np.concatenate([used[:, order], np.zeros((27, nt * 128 - n), bool)], 1).reshape(27, nt, 128).any(2)
    return float(u.sum(0).mean()), int(u.sum())


def digest_key(mask):
    """15-bit sort key that groups rows by neighbourhood STRUCTURE instead of by the numeric value of
    the mask (offset k = (dz+1)*9 + (dy+1)*3 + (dx+1)): for the plane below and the plane above, one bit
    per dy row ("any neighbour in that row"), then the nine bits of the voxel's own z plane.  LiDAR
    surfaces are mostly thin sheets: most rows have no neighbour above / below at all."""
    lower, centre, upper = mask & 0x1FF, (mask >> 9) & 0x1FF, (mask >> 18) & 0x1FF
    key = centre.copy()
    for plane, base in ((upper, 9), (lower, 12)):
        for row in range(3):
            key |= (((plane >> (3 * row)) & 7) != 0).astype(np.int64) << (base + row)
    return key


def main():
    for prof, sweeps in (('S', 1), ('L', 10)):
        pts = synthetic.lidar_scene(seed=0, sweeps=sweeps)
        _, c, _ = cpu.hard_voxelize(pts, synthetic.VOXEL_SIZE, synthetic.POINT_CLOUD_RANGE, 10, 160000)
        idx = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1)
        shape = [41, 1440, 1440]
        for level in range(4):
            used = cpu.subm_rulebook(idx, shape, 3, 1) >= 0           # (27, N)
            mask = (used.astype(np.int64) * (1 << np.arange(27))[:, None]).sum(0)
            base = tiles_active(used, np.arange(idx.shape[0]))
            srt = tiles_active(used, np.argsort(mask, kind='stable'))
            dig = tiles_active(used, np.argsort(digest_key(mask), kind='stable'))
            print('profile %s level %d: N %6d, pair density %.3f, active offsets per 128-row tile: %.1f now -> %.1f '
                  'sorted by the 27-bit mask (x%.2f) -> %.1f sorted by the 15-bit digest (x%.2f)'
                  % (prof, level, idx.shape[0], used.mean(), base[0], srt[0], base[1] / srt[1], dig[0],
                     base[1] / dig[1]))
            if level < 3:
                idx, _, shape = cpu.conv_rulebook(idx, shape, 3, 2, 1 if level < 2 else (0, 1, 1), 1)


if __name__ == '__main__':
    main()
