"""Hand-derived known-answer vectors for the rec0 path.

Every expected value below was worked out BY HAND from the cited reference lines (not by running the oracle, the CUDA path or
the Python model), like SURVEY.md Appendix C: the oracle (CPU suite) and the CUDA path (GPU suite) are both checked against
them, so these are pins of the restatement, not regression locks.  Patch tuple layout = abi.PATCH_DTYPE:
(u0, v0, size_u0, size_v0, u1, v1, d1, lod_x, lod_y, normal, tangent, bitangent, mode, orientation, additional_plane, pad).
"""
import numpy as np

from tmc2rs_b200 import abi

ROT90 = 2


def _gof(W, H, occ, geo0, geo1, patches, **params):
    geo = np.stack([geo0, geo1])[None].astype(np.uint16)
    ay = np.full((1, 2, H, W), 512, np.uint16)
    au = np.full((1, 2, H // 2, W // 2), 512, np.uint16)
    return abi.Gof(W, H, occ[None].astype(np.uint8), geo, ay, au, au.copy(), [np.array(patches, dtype=abi.PATCH_DTYPE)],
                   abi.Params(**params))


def rot90_reference_quirk():
    """A Rot90 patch in the reference's own pixel mapping (src/decoder.rs:853-866): size_uv0 stays in BLOCKS at pixel level.

    48x32 atlas (blocks 3x2), one patch uv0 = (1, 0), size_uv0 = (1, 1), Rot90, projection 0 (axes n,t,b = 0,2,1, mode 0),
    uv1 = (0, 0), d1 = 0.
      block level   (resolution 1):  (x, y) = (size_v0 - 1 - v + u0, u + v0) = (1 - 1 - 0 + 1, 0) = (1, 0)  -> block index 1
      pixel level   (resolution 16): (x, y) = (size_v0 - 1 - v + 16, u + 0)  = (16 - v, u)                  [size_v0 = 1, unscaled]
    so the patch's 256 pixels are canvas columns 1..16, rows 0..15 -- mostly canvas block 0 -- while it owns block 1.
    Occupancy: low-resolution sample (0,0) = 1, i.e. canvas x,y in 0..3.  The patch sees it at x = 16 - v in {1,2,3}
    (v = 13,14,15) and y = u in 0..3: non_zero_pixel > 0 -> block_to_patch[1] = 1 (codec.rs:242), block 0 stays 0.
    Geometry: depth0 = x + 2y (sample 4*(x+2y)+1), depth1 = x + 2y + 1 (sample 4*(x+2y+1)): two points per pixel.
    generate_point (decoder.rs:871-878): point = (depth + 0, v, u).  Loop order v outer, u inner (codec.rs:382-385).
    """
    W, H = 48, 32
    occ = np.zeros((H // 4, W // 4))
    occ[0, 0] = 1
    Y, X = np.mgrid[0:H, 0:W]
    g = _gof(W, H, occ, 4 * (X + 2 * Y) + 1, 4 * (X + 2 * Y + 1),
             [(1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 2, 1, 0, ROT90, 0, (0, 0))])
    pos, pix = [], []
    for v in (13, 14, 15):
        for u in (0, 1, 2, 3):
            x, y = 16 - v, u
            pos += [[x + 2 * y, v, u], [x + 2 * y + 1, v, u]]
            pix += [[x, y, 0], [x, y, 1]]
    want = {"block_to_patch": [0, 1, 0, 0, 0, 0], "positions": pos, "point_to_pixel": pix, "partition": [0] * 24}
    # spot values written out: first pixel v=13,u=0 -> canvas (3,0): depths 3 and 4
    assert pos[0] == [3, 13, 0] and pos[1] == [4, 13, 0] and pos[2] == [5, 13, 1] and pos[8] == [2, 14, 0] and pos[23] == [8, 15, 3]
    return g, want


def d1_wrap_and_mode1_clamp(absolute_d1):
    """Differential D1 in wrapping u16 (src/codec.rs:551-558), `as u16` truncation (decoder.rs:874) and the mode-1 clamp
    max(d1, depth) - depth (decoder.rs:885).

    32x16 atlas, two Default patches of one block each.
      patch 0: uv0 (0,0), projection 0 (n,t,b = 0,2,1; mode 0), d1 = 65530, uv1 = (1, 2)
      patch 1: uv0 (1,0), projection 3 (n,t,b = 0,2,1; mode 1), d1 = 5,     uv1 = (0, 0)
    Occupancy: samples (0,0) = 7 and (4,0) = 9 -> canvas x in 0..3 / 16..19, y in 0..3 (any non-zero value counts, codec.rs:393).
    Depths (sample / 4): block 0: depth0 = 2, depth1 = 10.  block 1: depth0 = 3 on even rows, 9 on odd rows; depth1 = 7.
      patch 0, pixel (u,v):  p0 = ((2 + 65530) as u16, v + 2, u + 1) = (65532, v+2, u+1)
          differential: p1[0] = 65532 + 10 (u16, wraps) = 6          absolute: p1[0] = (10 + 65530) as u16 = 4
      patch 1, even v:       p0 = (max(5,3) - 3, v, u) = (2, v, u)
          differential: p1[0] = 2 - 7 (u16, wraps) = 65531           absolute: p1[0] = max(5,7) - 7 = 0
      patch 1, odd v:        p0 = (max(5,9) - 9, v, u) = (0, v, u)
          differential: p1[0] = 0 - 7 = 65529                        absolute: p1 = (0, v, u) == p0 -> skipped (codec.rs:425)
    """
    W, H = 32, 16
    occ = np.zeros((H // 4, W // 4))
    occ[0, 0], occ[0, 4] = 7, 9
    Y, X = np.mgrid[0:H, 0:W]
    d0 = np.where(X < 16, 2, np.where(Y % 2 == 0, 3, 9))
    d1 = np.where(X < 16, 10, 7)
    g = _gof(W, H, occ, 4 * d0 + 3, 4 * d1 + 2,
             [(0, 0, 1, 1, 1, 2, 65530, 1, 1, 0, 2, 1, 0, 0, 0, (0, 0)), (1, 0, 1, 1, 0, 0, 5, 1, 1, 0, 2, 1, 1, 0, 0, (0, 0))],
             absolute_d1=absolute_d1)
    pos, part = [], []
    for v in range(4):
        for u in range(4):
            pos += [[65532, v + 2, u + 1], [4 if absolute_d1 else 6, v + 2, u + 1]]
            part += [0, 0]
    for v in range(4):
        for u in range(4):
            if v % 2 == 0:
                pos += [[2, v, u], [0 if absolute_d1 else 65531, v, u]]
                part += [1, 1]
            elif absolute_d1:
                pos += [[0, v, u]]
                part += [1]
            else:
                pos += [[0, v, u], [65529, v, u]]
                part += [1, 1]
    assert len(pos) == (56 if absolute_d1 else 64)
    return g, {"block_to_patch": [1, 2], "positions": pos, "partition": part}


def overlapping_patches_precedence():
    """Later patches overwrite earlier ones in block_to_patch (src/codec.rs:242-244); a patch only emits the blocks it still
    owns (codec.rs:379).

    32x32 atlas (blocks 2x2).  patch 0: uv0 (0,0), size (2,1): blocks 0 and 1.  patch 1: uv0 (1,0), size (1,2): blocks 1 and 3.
    patch 2: uv0 (0,1), size (2,1): blocks 2 and 3 -- but block 3 has no occupancy anywhere, so nobody owns it.
    Occupancy: one sample per block 0, 1, 2: canvas pixels (0..3, 0..3), (16..19, 0..3), (0..3, 16..19).
    block_to_patch = [1, 2, 3, 0].  Depth 1 in both maps everywhere (duplicates: one point per pixel), projection 2
    (n,t,b = 2,0,1; mode 0), d1 = 0: point = (u + u1, v + v1, 1) with uv1 = (100,0) / (200,0) / (300,0) per patch.
      patch 0 emits block 0 only: pixels u,v in 0..3          -> (100 + u, v, 1)
      patch 1 emits block 1:      local u,v in 0..3           -> (200 + u, v, 1)      (its block (0,1) = canvas block 3: unowned)
      patch 2 emits block 2:      local u in 0..3, v in 0..3  -> (300 + u, v, 1)
    """
    W, H = 32, 32
    occ = np.zeros((H // 4, W // 4))
    occ[0, 0] = occ[0, 4] = occ[4, 0] = 255
    ones = np.full((H, W), 4 * 1 + 1)
    g = _gof(W, H, occ, ones, ones + 1,
             [(0, 0, 2, 1, 100, 0, 0, 1, 1, 2, 0, 1, 0, 0, 0, (0, 0)), (1, 0, 1, 2, 200, 0, 0, 1, 1, 2, 0, 1, 0, 0, 0, (0, 0)),
              (0, 1, 2, 1, 300, 0, 0, 1, 1, 2, 0, 1, 0, 0, 0, (0, 0))])
    pos, part = [], []
    for k, base in enumerate((100, 200, 300)):
        for v in range(4):
            for u in range(4):
                pos.append([base + u, v, 1])
                part.append(k)
    return g, {"block_to_patch": [1, 2, 3, 0], "positions": pos, "partition": part}


ALL = [("rot90_reference_quirk", rot90_reference_quirk), ("d1_differential", lambda: d1_wrap_and_mode1_clamp(False)),
       ("d1_absolute", lambda: d1_wrap_and_mode1_clamp(True)), ("precedence", overlapping_patches_precedence)]


def check(result, want, what):
    assert result["block_to_patch"].tolist() == want["block_to_patch"], what
    assert result["point_count"] == len(want["positions"]), (what, result["point_count"])
    assert result["positions"].tolist() == want["positions"], what
    assert result["partition"].tolist() == want["partition"], what
    if "point_to_pixel" in want:
        assert result["point_to_pixel"].tolist() == want["point_to_pixel"], what
    assert (np.asarray(result["colors"]) == 127).all(), what          # Y = U = V = 512 -> (127, 127, 127), codec.rs:661-687
