"""Host logic (no GPU): synthetic generator determinism, GOF views, golden fixtures, frame sharding."""
import hashlib
import os

import numpy as np

import tmc2rs_b200  # noqa: F401
from oracle import oracle
from tmc2rs_b200 import abi, shard, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_synth_is_deterministic_and_shaped():
    cfg = synth.config("small")
    a, b = synth.make_gof(cfg), synth.make_gof(cfg)
    assert digest(a.occ, a.geo, a.attr_y, a.attr_u, a.attr_v) == digest(b.occ, b.geo, b.attr_y, b.attr_u, b.attr_v)
    assert all(np.array_equal(p, q) for p, q in zip(a.patches, b.patches))
    assert a.occ.shape == (3, 64, 64) and a.geo.shape == (3, 2, 256, 256) and a.attr_u.shape == (3, 2, 128, 128)
    assert a.geo.max() < 1024 and a.attr_y.max() <= 1023
    # occupancy uses "non-zero", not "== 1"
    vals = np.unique(a.occ)
    assert 1 in vals and 255 in vals and len(vals) > 10


def test_synth_c1_shape_matches_baseline_config():
    g = synth.make_gof(synth.config("c1"))
    assert (g.width, g.height, g.frame_count) == (1024, 1024, 1)
    r = oracle.reconstruct_frame(abi.GofView(g), 0)
    assert 700_000 < r["point_count"] < 900_000          # "~800k pts"
    # overlapping patches exercise block-to-patch precedence: some block is covered by >= 2 patches
    cover = np.zeros((64, 64), int)
    for p in g.patches[0]:
        w, h = (p["size_u0"], p["size_v0"]) if p["patch_orientation"] == 0 else (p["size_v0"], p["size_u0"])
        cover[p["v0"]:p["v0"] + h, p["u0"]:p["u0"] + w] += 1
    assert cover.max() >= 2


def test_golden_fixture_small():
    """tests/golden/small_f0.npz was written by tests/golden/make_golden.py from the oracle; the Appendix-C vector and the
    colour KATs (hand-derived from the reference text) are checked in test_oracle_kat.py."""
    z = np.load(os.path.join(GOLD, "small_f0.npz"))
    g = synth.make_gof(synth.config("small"))
    r = oracle.reconstruct_frame(abi.GofView(g), 0)
    assert int(z["point_count"]) == r["point_count"]
    assert z["sha_positions"].item() == digest(r["positions"])
    assert z["sha_colors"].item() == digest(r["colors"])
    assert np.array_equal(z["block_to_patch"], r["block_to_patch"])
    assert np.array_equal(z["positions_head"], r["positions"][:256])
    assert np.array_equal(z["colors_head"], r["colors"][:256])


def _load_make_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_golden_fixture_smoothing():
    """tests/golden/small_smooth_f1.npz freezes this repository's own smoothing specification (boundary classes, grid geometry
    smoothing, grid colour smoothing; the reference has only stubs): the oracle must still produce it."""
    z = np.load(os.path.join(GOLD, "small_smooth_f1.npz"))
    r = oracle.reconstruct_frame(abi.GofView(_load_make_golden().smoothing_case()), 1)
    assert int(z["point_count"]) == r["point_count"]
    assert int(z["smoothed_positions"]) == r["smoothed_positions"] and int(z["smoothed_colors"]) == r["smoothed_colors"]
    assert z["sha_boundary_type"].item() == digest(r["boundary_type"])
    assert z["sha_positions"].item() == digest(r["positions"])
    assert z["sha_colors"].item() == digest(r["colors"])
    assert np.array_equal(z["moved_positions_head"], r["positions"][z["moved_index_head"]])


def test_frame_sharding_covers_every_frame_once():
    for total in (0, 1, 7, 32, 300):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.frames_for_rank(total, r, world)
                seen += list(range(lo, hi))
                assert hi - lo in (total // world, total // world + 1)
            assert seen == list(range(total))


def test_gof_view_keeps_pointers_and_strides():
    g = synth.make_gof(synth.config("tiny"))
    v = abi.GofView(g)
    assert v.c.frame_count == 2 and v.c.geo_video_frames == 4 and v.c.attr_video_frames == 4
    assert v.c.frames[1].geo[1] == g.geo[1, 1].ctypes.data
    assert v.c.frames[0].attr_stride_c == g.width // 2
    assert v.c.frames[0].patch_count == len(g.patches[0])


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 37                                            # frames of a sequence, sharded frame-wise
        lo, hi = shard.frames_for_rank(total, rank, world)
        # every rank "reconstructs" its slice: elapsed differs per rank, points = a known function of the frame index
        pts = sum(1000 + f for f in range(lo, hi))
        ms, p, fr = shard.reduce_metrics(10.0 * (rank + 1), pts, hi - lo, "cpu")
        q.put((rank, lo, hi, ms, p, fr))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_metric_reduction():
    """N>1 host logic on CPU: world_size 2 over gloo -- disjoint frame slices, max-over-ranks time, summed counts; no
    data-path collective exists on this path."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, lo0, hi0, ms0, p0, f0), (_, lo1, hi1, ms1, p1, f1) = res
    assert (lo0, hi1) == (0, 37) and hi0 == lo1                    # contiguous, disjoint, complete
    assert ms0 == ms1 == 20.0                                      # max over ranks
    assert p0 == p1 == sum(1000 + f for f in range(37)) and f0 == f1 == 37


def test_ply_writer_matches_the_reference_format():
    """tmc2rs_b200/ply.py against src/writer.rs:31-75 written out by hand for a two-point cloud."""
    from tmc2rs_b200 import ply
    pos = np.array([[1, 2, 3], [65535, 0, 7]], np.uint16)
    col = np.array([[255, 0, 127], [1, 2, 3]], np.uint8)
    want = ("ply\nformat ascii 1.0\nelement vertex 2\nproperty uint x\nproperty uint y\nproperty uint z\n"
            "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 0\n"
            "property list uint8 int32 vertex_index\nend_header\n1 2 3 255 0 127\n65535 0 7 1 2 3\n")
    assert ply.ascii_ply(pos, col) == want.encode()
    assert ply.ascii_ply(pos, None).endswith(b"end_header\n1 2 3\n65535 0 7\n") and b"red" not in ply.ascii_ply(pos, None)
    for data in (ply.ascii_ply(pos, col), ply.binary_ply(pos, col)):
        p2, c2 = ply.read_ply(data)
        assert np.array_equal(p2, pos) and np.array_equal(c2, col)
    assert len(ply.binary_ply(pos, col)) == len(ply._header("binary_little_endian", 2, True)) + 2 * 15


def test_compare_dump_report(tmp_path):
    """tools/compare_dump.py: the +-1 / differing-count report of north_star on synthetic pairs."""
    import importlib.util
    from tmc2rs_b200 import ply
    spec = importlib.util.spec_from_file_location("compare_dump", os.path.join(os.path.dirname(GOLD), "..", "tools", "compare_dump.py"))
    cd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cd)
    rng = np.random.RandomState(3)
    pos = rng.randint(0, 1024, (500, 3)).astype(np.uint16)
    col = rng.randint(0, 256, (500, 3)).astype(np.uint8)
    pos2, col2 = pos.copy(), col.copy()
    pos2[10, 1] += 1; pos2[20, 2] += 3; col2[5, 0] ^= 1                       # two moved points (one beyond +-1), one recoloured
    (tmp_path / "ref.ply").write_bytes(ply.ascii_ply(pos, col))
    (tmp_path / "ours.ply").write_bytes(ply.binary_ply(pos2, col2))
    pos2.astype("<u2").tofile(tmp_path / "ours.u16"); col2.tofile(tmp_path / "ours.u8")
    for ours in ("ours.ply", "ours.u16"):
        rep = cd.compare(cd.load(str(tmp_path / "ref.ply")), cd.load(str(tmp_path / ours)))
        assert rep["positions_differing"] == 2 and rep["positions_beyond_tolerance"] == 1 and rep["colors_differing"] == 1
        assert rep["colors_beyond_tolerance"] == 0 and not rep["within_tolerance"] and not rep["bit_exact"]
    same = cd.compare(cd.load(str(tmp_path / "ref.ply")), cd.load(str(tmp_path / "ref.ply")))
    assert same["bit_exact"]
    perm = rng.permutation(500)
    (tmp_path / "shuffled.ply").write_bytes(ply.ascii_ply(pos[perm], col[perm]))
    assert cd.compare(cd.load(str(tmp_path / "ref.ply")), cd.load(str(tmp_path / "shuffled.ply")), unordered=True)["bit_exact"]
    assert cd.main([str(tmp_path / "ref.ply"), str(tmp_path / "ref.ply"), "--json"]) == 0
