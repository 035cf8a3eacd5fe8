"""Independent pure-Python model of the reference's reconstruction (small cases only).

Written directly from the reference text (tmc2-rs src/codec.rs:205-250, :288-300, :352-480, :517-565, :569-658, :661-687;
src/decoder.rs:853-888, :973-980) WITHOUT looking at oracle/tmc2_oracle.c's structure: plain loops, Python integers with
explicit 64-bit / 16-bit wrapping.  It is the second opinion that pins the C oracle (the reference itself cannot be built
in this image: no cargo/rustc).  Test infrastructure only.
"""
import math

M64 = (1 << 64) - 1


class Panic(Exception):
    """The reference would panic (assert!/unwrap/unimplemented!)."""


def to_canvas(p, u, v, resolution, size_scale):
    """decoder.rs:853-867 with usize wrapping."""
    u0, v0 = p["u0"] * resolution, p["v0"] * resolution
    su, sv = p["size_u0"] * size_scale, p["size_v0"] * size_scale
    o = p["patch_orientation"]
    if o == 0:
        x, y = u + u0, v + v0
    elif o == 2:
        x, y = sv - 1 - v + u0, u + v0
    elif o == 3:
        x, y = su - 1 - u + u0, sv - 1 - v + v0
    elif o == 4:
        x, y = v + u0, su - 1 - u + v0
    elif o == 5:
        x, y = su - 1 - u + u0, v + v0
    elif o == 6:
        x, y = sv - 1 - v + u0, su - 1 - u + v0
    elif o == 7:
        x, y = u + u0, sv - 1 - v + v0
    else:  # Swap (1), MRot270 (8)
        x, y = v + u0, u + v0
    return x & M64, y & M64


def gen_point(p, u, v, depth):
    """decoder.rs:871-888."""
    pt = [0, 0, 0]
    if p["projection_mode"] == 0:
        n = depth + p["d1"]
    else:
        n = max(p["d1"], depth) - depth
    pt[p["normal_axis"]] = n & 0xFFFF
    pt[p["tangent_axis"]] = (u * p["lod_x"] + p["u1"]) & 0xFFFF
    pt[p["bitangent_axis"]] = (v * p["lod_y"] + p["v1"]) & 0xFFFF
    return pt


def yuv_to_rgb(c):
    """codec.rs:661-687 (Python floats are IEEE doubles; every operation rounds separately)."""
    y, u, v = float(c[0]), float(c[1]), float(c[2])

    def clamp(x):
        if x < 0.0:
            return 0
        if x > 255.0:
            return 255
        return int(x)
    r = y + 1.57480 * (v - 512.0)
    g = y - 0.18733 * (u - 512.0) - (0.46813 * (v - 512.0))
    b = y + 1.85563 * (u - 512.0)
    return [clamp(math.floor(r / 1023.0 * 255.0)), clamp(math.floor(g / 1023.0 * 255.0)),
            clamp(math.floor(b / 1023.0 * 255.0))]


def reconstruct(gof, f):
    """Returns dict(block_to_patch, occupancy_map, positions, colors16bit, colors, partition, point_to_pixel)."""
    P = gof.params
    W, H = gof.width, gof.height
    res, prec = P.occupancy_resolution, P.occupancy_precision
    sscale = res if P.orientation_mode == 1 else 1
    occ = gof.occ[f]
    oh, ow = occ.shape
    patches = [{k: int(p[k]) for k in p.dtype.names if k != "_reserved"} for p in gof.patches[f]]
    bw, bh = W // res, H // res

    def occ_get(x, y):
        if not (x < ow and y < oh):
            raise Panic("occupancy index")
        return int(occ[y, x])

    # codec.rs:205-250
    b2p = [0] * (bw * bh)
    for pi, p in enumerate(patches):
        for v0 in range(p["size_v0"]):
            for u0 in range(p["size_u0"]):
                bx, by = to_canvas(p, u0, v0, 1, 1)
                if not (bx < bw and by < bh):
                    raise Panic("block outside canvas")
                nz = 0
                for v1 in range(res):
                    for u1 in range(res):
                        x, y = to_canvas(p, u0 * res + u1, v0 * res + v1, res, sscale)
                        if not (x < W and y < H):
                            raise Panic("pixel outside canvas")
                        nz += occ_get(x // prec, y // prec)
                if nz > 0:
                    b2p[by * bw + bx] = pi + 1
    # codec.rs:288-300
    occ_full = [[occ_get(u // prec, v // prec) for u in range(W)] for v in range(H)]
    positions, partition, ptp = [], [], []
    for pi, p in enumerate(patches):
        for v0 in range(p["size_v0"]):
            for u0 in range(p["size_u0"]):
                bx, by = to_canvas(p, u0, v0, 1, 1)
                if b2p[by * bw + bx] != pi + 1:
                    continue
                for v1 in range(res):
                    for u1 in range(res):
                        u, v = u0 * res + u1, v0 * res + v1
                        x, y = to_canvas(p, u, v, res, sscale)
                        if not (x < W and y < H):
                            raise Panic("pixel outside canvas")
                        if occ_full[y][x] == 0:
                            continue
                        d0 = int(gof.geo[f, 0, y, x]) // 4
                        d1 = int(gof.geo[f, 1, y, x]) // 4
                        p0 = gen_point(p, u, v, d0)
                        if P.absolute_d1:
                            p1 = gen_point(p, u, v, d1)
                        else:
                            p1 = list(p0)
                            a = p["normal_axis"]
                            p1[a] = (p1[a] + d1) & 0xFFFF if p["projection_mode"] == 0 else (p1[a] - d1) & 0xFFFF
                        for i, pt in enumerate((p0, p1)):
                            if i != 0 and pt == p0:
                                continue
                            positions.append(pt)
                            partition.append(pi)
                            ptp.append([x, y, i])
    c16, rgb = [], []
    if P.attribute_count:
        for (x, y, z) in ptp:
            c = [int(gof.attr_y[f, z, y, x]), int(gof.attr_u[f, z, y // 2, x // 2]), int(gof.attr_v[f, z, y // 2, x // 2])]
            c16.append(c)
            rgb.append(yuv_to_rgb(c))
    return {"block_to_patch": b2p, "occupancy_map": occ_full, "positions": positions, "colors16bit": c16,
            "colors": rgb, "partition": partition, "point_to_pixel": ptp}
