"""The reference-shaped renderer class end to end on the GPU, plus size-independent properties at
sizes the oracle cannot reach (BASELINE.json full resolution)."""
import os

import numpy as np
import pytest
import torch

import golden_fixtures
from conftest import GOLDEN_SCENES

pytestmark = pytest.mark.gpu


def _renderer(name="wavefront_comparison", width=48, aspect=16 / 9, spp=4, **kw):
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.renderer_factory import RendererFactory
    from path_tracer_python_b200.scene import camera, point3, vec3
    cs, meta = golden_fixtures.load_config(name, width, aspect)
    cam = camera()
    cam.aspect_ratio, cam.img_width, cam.samples_per_pixel = aspect, width, spp
    c = scenes._CAMERAS[name]
    cam.vfov, cam.lookfrom, cam.lookat, cam.vup = c["vfov"], point3(*c["lookfrom"]), point3(*c["lookat"]), vec3(*c["vup"])
    r = RendererFactory.create("taichi", None, cam, kw.pop("path", "/tmp/pt_test.png"), compiled_scene=cs, quiet=True, **kw)
    return r, cs, meta


def test_render_call_matches_reference_golden(tmp_path):
    """RendererFactory.create('taichi', ...) -> set knobs -> render(): the reference's call sequence
    (src/scenes.py:1504-1514) reproduces the reference's accumulation buffer bit for bit."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    _, gmeta, out = golden_fixtures.load_fixture("wavefront_comparison")
    path = str(tmp_path / "mega.png")
    r, cs, meta = _renderer(spp=gmeta["spp"], seed=gmeta["rng_seed"], path=path)
    r.background_color = gmeta["background"]
    r.max_depth = 50
    r.render(enable_preview=False)
    assert np.array_equal(r.accum_numpy(), out["accum_megakernel"])
    assert len(r.sample_times) == gmeta["spp"] and os.path.getsize(path) > 100
    from PIL import Image
    img = np.array(Image.open(path))
    from path_tracer_python_b200.preview import buffer_to_image
    assert np.array_equal(img, buffer_to_image(out["accum_megakernel"], gmeta["spp"]))   # device tonemap == preview.py
    r2, _, _ = _renderer(spp=gmeta["spp"], seed=gmeta["rng_seed"], path=str(tmp_path / "wave.png"))
    r2.background_color = gmeta["background"]
    r2.render_wavefront(enable_preview=False)
    assert np.array_equal(r2.accum_numpy(), out["accum_wavefront"])


def test_interactive_viewer_compat_surface():
    r, _, _ = _renderer(spp=3, seed=5)
    r.background_color = (0.7, 0.8, 1.0)
    r._upload_camera_to_gpu()
    r.clear_accumulation_buffer()
    for s in range(3):
        r.render_sample(s)
    r._sync_gpu_to_cpu()
    a = r.accum_numpy().copy()
    r.clear_accumulation_buffer()
    assert r.accum_numpy().max() == 0 and r.current_sample == 0
    r.engine.render_megakernel(r._accum, 0, 3, 5)
    assert np.array_equal(r.accum_numpy(), a)            # 3 x render_sample == one 3-sample launch
    assert r.num_spheres == 41 and r.num_bvh_nodes == 81
    assert 1.0 < r._get_average_depth() < 10.0
    for k in ("scene_compile", "bvh_compile", "gpu_upload", "total_setup"):
        assert k in r.timing


def test_full_resolution_properties():
    """Config 2 at full size (1920x1080): determinism, sample-range and row-range composition
    (the multi-GPU splits) are bit-exact; megakernel and wavefront agree statistically."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.engine import Engine
    e = Engine(0)
    cs, meta = golden_fixtures.load_config("vol2_final_scene", 1920, 16 / 9)
    e.set_scene(cs)
    e.set_camera(meta["camera"], (0, 0, 0), 50)
    W, H = 1920, meta["height"]
    a = e.new_accum(W, H); e.render_megakernel(a, 0, 4, 11)
    b = e.new_accum(W, H); e.render_megakernel(b, 0, 4, 11)
    assert torch.equal(a, b)
    c = e.new_accum(W, H)
    e.render_megakernel(c, 0, 1, 11); e.render_megakernel(c, 1, 3, 11)
    assert torch.equal(a, c)
    d = e.new_accum(W, H)
    e.render_megakernel(d, 0, 4, 11, y0=0, y1=540); e.render_megakernel(d, 0, 4, 11, y0=540, y1=H)
    assert torch.equal(a, d)
    w = e.new_accum(W, H); e.render_wavefront(w, 0, 4, 11)
    w2 = e.new_accum(W, H); e.render_wavefront(w2, 0, 2, 11); e.render_wavefront(w2, 2, 2, 11)
    assert torch.equal(w, w2)
    assert torch.isfinite(a).all() and torch.isfinite(w).all() and (a >= 0).all()
    ma, mw = a.mean().item() / 4, w.mean().item() / 4
    assert abs(ma - mw) / ma < 0.01                      # same estimator, different primary-ray normalisation
    # different seeds decorrelate; same seed different sample ranges differ
    s2 = e.new_accum(W, H); e.render_megakernel(s2, 0, 4, 12)
    assert not torch.equal(a, s2) and abs(s2.mean().item() / 4 - ma) / ma < 0.01


def test_edge_cases_empty_single_and_ragged():
    from path_tracer_python_b200.engine import Engine
    from path_tracer_python_b200 import scenes, bvh
    import golden_fixtures
    e = Engine(0)
    cam, h = scenes.camera_vectors((0, 0, 5), (0, 0, 0), (0, 1, 0), 40, 37, 37 / 13)   # ragged 37x13: not a tile multiple
    # empty scene: every path returns the background
    e.set_scene({})
    e.set_camera(cam, (0.25, 0.5, 0.75), 50)
    for fn in (e.render_megakernel, e.render_wavefront):
        acc = e.new_accum(37, h); fn(acc, 0, 3, 1)
        got = acc.cpu().numpy()
        assert np.array_equal(got, np.broadcast_to(np.float32([0.75, 1.5, 2.25]), got.shape))
    # one emissive sphere: root is a leaf
    one = {"sphere_data": np.float32([[0, 0, 0, 1]]), "sphere_material_type": np.int32([3]),
           "sphere_material_emit_color": np.float32([[2, 3, 4]])}
    for k in ("material_albedo", "texture_color1", "texture_color2", "medium_albedo"):
        one["sphere_" + k] = np.zeros((1, 3), np.float32)
    for k in ("material_fuzz", "material_ir", "texture_scale", "medium_density"):
        one["sphere_" + k] = np.zeros(1, np.float32)
    for k in ("texture_type", "texture_image_idx", "is_constant_medium"):
        one["sphere_" + k] = np.zeros(1, np.int32)
    one.update(bvh.compile_bvh(one))
    e.set_scene(one)
    e.set_camera(cam, (0, 0, 0), 50)
    acc = e.new_accum(37, h); e.render_megakernel(acc, 0, 2, 1)
    got = acc.cpu().numpy()
    centre = got[h // 2, 18]
    assert np.array_equal(centre, np.float32([4, 6, 8])) and np.array_equal(got[0, 0], np.float32([0, 0, 0]))
    wf = e.new_accum(37, h); e.render_wavefront(wf, 0, 2, 1)
    assert np.array_equal(wf.cpu().numpy() > 0, got > 0)
    # zero samples / empty row range are no-ops; bad arguments raise
    z = e.new_accum(37, h); e.render_megakernel(z, 0, 0, 1); e.render_megakernel(z, 0, 2, 1, y0=5, y1=5)
    assert z.max().item() == 0
    from path_tracer_python_b200._lib import B200PTError
    with pytest.raises(B200PTError):
        e.render_megakernel(z, 0, 1, 1, y0=3, y1=h + 1)


def test_counters_match_oracle_semantics(oracle_mod):
    """paths/bounces/rays are traversal-order independent and must equal the oracle's exactly;
    box tests differ (different visiting order) but stay within 2x of the reference-order count."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.engine import Engine
    O = oracle_mod
    cs, meta, _ = golden_fixtures.load_fixture("vol2_final_scene")
    e = Engine(0)
    e.set_scene(cs); e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    e.enable_counters(True); e.counters(reset=True)
    acc = e.new_accum(meta["width"], meta["height"]); e.render_megakernel(acc, 0, meta["spp"], meta["rng_seed"])
    g = e.counters()
    S = O.OracleScene(cs); cam = O.make_camera(meta["camera"], meta["background"], meta["max_depth"])
    _, c = O.render(S, cam, meta["width"], meta["height"], meta["spp"], seed=meta["rng_seed"])
    assert g["paths"] == c["paths"] and g["bounces"] == c["bounces"] and g["rays"] == c["rays"]
    assert 0.3 < g["box_tests"] / c["node_tests"] < 2.0


@pytest.mark.parametrize("scene_name", ["feature_zoo", "vol2_final_scene"])
def test_results_do_not_depend_on_tuning_options(scene_name):
    """Scheduling knobs (burst / vote thresholds, shared-memory nodes, block size, wavefront batch
    size, number of full waves, fused leading waves, material sort) change how the work is laid out
    on the machine, never the result: every combination reproduces the reference bit for bit."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.engine import Engine
    cs, meta, out = golden_fixtures.load_fixture(scene_name)
    e = Engine(0)
    e.set_scene(cs)
    e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H, spp, seed = meta["width"], meta["height"], meta["spp"], meta["rng_seed"]
    mega_opts = [dict(burst_num=0, vote_num=0), dict(burst_num=32, vote_num=32), dict(burst_num=7, vote_num=3, smem_nodes=0),
                 dict(block_threads=64, smem_nodes=9), dict(block_threads=128, sync_samples=1), dict(sync_samples=3, smem_nodes=100000)]
    for o in mega_opts:
        for k, v in dict(dict(burst_num=12, vote_num=12, smem_nodes=-1, block_threads=256, sync_samples=0), **o).items():
            e.set_option(k, v)
        acc = e.new_accum(W, H)
        e.render_megakernel(acc, 0, spp, seed)
        assert np.array_equal(acc.cpu().numpy(), out["accum_megakernel"]), o
    wf_opts = [dict(wf_waves=1), dict(wf_waves=2, wf_fuse_primary=0), dict(wf_waves=50, wf_fuse_primary=0, wavefront_sort=0),
               dict(wf_fuse_primary=3, wf_waves=4), dict(wf_fuse_primary=7, wf_waves=3), dict(wf_burst_num=0, wf_vote_num=0, wf_blocks_per_sm=1),
               dict(wf_burst_num=32, wf_vote_num=32, wf_blocks_per_sm=8, smem_nodes=0), dict(wf_tile=0, wf_extend_minb=5),
               dict(wf_extend_minb=6, wf_blocks_per_sm=6, wf_grid_waves=1), dict(wf_grid_waves=4096, wf_waves=3), dict(wf_rel_primary=0), dict(wf_tail_vote=0, wf_waves=2),
               dict(wf_tail_vote=1, wf_waves=1, wf_burst_num=5, wf_vote_num=3), dict(wf_overlap=0), dict(wf_overlap=1, wf_waves=6),
               dict(wf_overlap=1, wf_waves=5), dict(wf_overlap=1, wf_fuse_primary=2, wf_waves=5),
               dict(wf_shade_split=0), dict(wf_shade_split=0, wf_waves=3, wf_overlap=0), dict(wf_shade_split=1, wf_fuse_primary=0, wf_waves=4),
               dict(wf_oct_nodes=0), dict(wf_oct_nodes=0, wf_rel_primary=0), dict(wf_oct_nodes=1, wf_rel_primary=0, wf_waves=3),
               dict(wf_extend_bulk=1), dict(wf_extend_bulk=1, wf_fuse_primary=0, wf_waves=50, wf_overlap=0),
               dict(wf_shade_stage=0), dict(wf_shade_stage=0, wf_shade_split=0), dict(wf_shade_stage=1, wf_shade_split=0, wf_overlap=0, wf_waves=3),
               dict(wf_shade_stage=1, wf_shade_split=0, wavefront_sort=0, wf_fuse_primary=0, wf_waves=50), dict(wf_extend_carveout=30)]
    for o in wf_opts:
        for k, v in dict(dict(wf_waves=12, wf_fuse_primary=1, wavefront_sort=1, wf_burst_num=20, wf_vote_num=16, wf_blocks_per_sm=6,
                              smem_nodes=-1, wf_tile=1, wf_extend_minb=4, wf_grid_waves=16, wf_rel_primary=1, wf_tail_vote=1, wf_overlap=1,
                              wf_shade_split=1, wf_oct_nodes=1, wf_extend_bulk=0, wf_shade_stage=1, wf_extend_carveout=-1), **o).items():
            e.set_option(k, v)
        for flight in (1, 3, 32):
            e.wavefront_samples_in_flight = flight
            acc = e.new_accum(W, H)
            e.render_wavefront(acc, 0, spp, seed)
            assert np.array_equal(acc.cpu().numpy(), out["accum_wavefront"]), (o, flight)


def test_stage_timing_reports_every_launch():
    """pt_enable_stage_timing / pt_get_stage_times: one record per kernel launch of the last wavefront
    call, in launch order, and switching it on does not change the image."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.engine import Engine
    cs, meta, out = golden_fixtures.load_fixture("feature_zoo")
    e = Engine(0)
    e.set_scene(cs)
    e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    W, H, spp, seed = meta["width"], meta["height"], meta["spp"], meta["rng_seed"]
    e.set_option("wf_waves", 4)
    e.wavefront_samples_in_flight = 2
    e.enable_stage_timing(True)
    acc = e.new_accum(W, H)
    e.render_wavefront(acc, 0, spp, seed)
    t = e.stage_times()
    assert np.array_equal(acc.cpu().numpy(), out["accum_wavefront"])
    batches = (spp + 1) // 2
    per_batch = ["reset", "primary"] + ["extend", "bin", "shade"] * 3 + ["tail", "resolve"]
    assert [x[0] for x in t] == per_batch * batches
    assert [x[1] for x in t if x[0] == "extend"] == [1, 2, 3] * batches
    assert all(x[2] >= 0.0 for x in t) and sum(x[2] for x in t) > 0.0
    e.enable_stage_timing(False)
    e.render_wavefront(acc, 0, 1, seed)
    assert e.stage_times() == []


def test_origin_relative_nodes_follow_camera_and_scene_changes():
    """The primary wave's origin-relative node array is a per-(scene, camera centre) cache inside the
    context: moving the camera or swapping the scene must rebuild it.  A context that renders camera A,
    then camera B, then another scene, gives the same images as fresh contexts do."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.engine import Engine
    cs, meta, _ = golden_fixtures.load_fixture("vol2_final_scene")
    cs2, meta2, _ = golden_fixtures.load_fixture("cornell_smoke")
    assert not meta["camera"]["cam_defocus_angle"] > 0          # the relative path is actually taken
    W, H, seed = meta["width"], meta["height"], meta["rng_seed"]
    cam_b = dict(meta["camera"])
    cam_b["cam_center"] = [c + d for c, d in zip(meta["camera"]["cam_center"], (13.0, -7.5, 40.0))]

    def fresh(scene, m, cam, w, h):
        e = Engine(0)
        e.set_scene(scene)
        e.set_camera(cam, m["background"], m["max_depth"])
        acc = e.new_accum(w, h)
        e.render_wavefront(acc, 0, 2, seed)
        return acc.cpu().numpy()

    e = Engine(0)
    e.set_scene(cs)
    e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
    a = e.new_accum(W, H); e.render_wavefront(a, 0, 2, seed)
    e.set_camera(cam_b, meta["background"], meta["max_depth"])
    b = e.new_accum(W, H); e.render_wavefront(b, 0, 2, seed)
    e.set_scene(cs2)
    e.set_camera(meta2["camera"], meta2["background"], meta2["max_depth"])
    c = e.new_accum(meta2["width"], meta2["height"]); e.render_wavefront(c, 0, 2, seed)
    assert np.array_equal(a.cpu().numpy(), fresh(cs, meta, meta["camera"], W, H))
    assert np.array_equal(b.cpu().numpy(), fresh(cs, meta, cam_b, W, H))
    assert not np.array_equal(a.cpu().numpy(), b.cpu().numpy())
    assert np.array_equal(c.cpu().numpy(), fresh(cs2, meta2, meta2["camera"], meta2["width"], meta2["height"]))
    e.set_option("wf_rel_primary", 0)
    c0 = e.new_accum(meta2["width"], meta2["height"]); e.render_wavefront(c0, 0, 2, seed)
    assert np.array_equal(c0.cpu().numpy(), c.cpu().numpy())


@pytest.mark.parametrize("mode", ["megakernel", "wavefront"])
def test_full_size_window_bit_exact_vs_oracle(oracle_mod, mode):
    """Config 2 at BASELINE size (1920x1080): rows 500..564 of the full-size frame, 4 samples, are
    bit-identical to the CPU oracle (a window the oracle finishes in seconds; the GPU renders it as
    a row range of the full frame, with the wavefront's defaults -- tiles, origin-relative primary
    nodes, 32 samples in flight)."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.engine import Engine
    O = oracle_mod
    cs, meta = golden_fixtures.load_config("vol2_final_scene", 1920, 16 / 9)
    W, H, y0, y1, spp, seed = 1920, meta["height"], 500, 564, 4, 2024
    e = Engine(0)
    e.set_scene(cs)
    e.set_camera(meta["camera"], (0.7, 0.8, 1.0), 50)
    acc = e.new_accum(W, H)
    (e.render_megakernel if mode == "megakernel" else e.render_wavefront)(acc, 0, spp, seed, y0=y0, y1=y1)
    got = acc.cpu().numpy()
    S = O.OracleScene(cs)
    cam = O.make_camera(meta["camera"], (0.7, 0.8, 1.0), 50)
    ref, _ = O.render(S, cam, W, H, spp, seed=seed, mode=mode, y0=y0, y1=y1)
    assert np.array_equal(got[y0:y1], ref[y0:y1])
    assert not got[:y0].any() and not got[y1:].any()


def test_one_context_many_scenes_with_batch_overlap():
    """One context renders all fixture scenes back to back, twice, with several batches per call (so
    the two-stream batch schedule and the per-call origin-relative node copy are exercised across
    scene / camera / workspace changes): every image equals the reference's."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.engine import Engine
    e = Engine(0)
    e.wavefront_samples_in_flight = 2
    for _ in range(2):
        for name in GOLDEN_SCENES:
            cs, meta, out = golden_fixtures.load_fixture(name)
            e.set_scene(cs)
            e.set_camera(meta["camera"], meta["background"], meta["max_depth"])
            acc = e.new_accum(meta["width"], meta["height"])
            e.render_wavefront(acc, 0, meta["spp"], meta["rng_seed"])
            assert np.array_equal(acc.cpu().numpy(), out["accum_wavefront"]), name


@pytest.mark.parametrize("mode", ["megakernel", "wavefront"])
def test_interleaved_row_windows_compose_to_the_full_frame(mode):
    """The multi-GPU tile split on one GPU: three 'ranks' render their interleaved row blocks (one call
    each, pt_set_row_interleave) into one buffer -- the union is bit-identical to the full-frame render.
    270 rows / 8-row blocks = 33.75 blocks: the last block sticks out of the image (rows skipped)."""
    from path_tracer_python_b200 import scenes
    import golden_fixtures
    from path_tracer_python_b200.distributed import interleaved_window
    from path_tracer_python_b200.engine import Engine
    e = Engine(0)
    cs, meta = golden_fixtures.load_config("vol2_final_scene", 480, 16 / 9)
    e.set_scene(cs)
    e.set_camera(meta["camera"], (0, 0, 0), 50)
    W, H, spp, seed = 480, meta["height"], 8, 3
    fn = e.render_megakernel if mode == "megakernel" else e.render_wavefront
    full = e.new_accum(W, H)
    fn(full, 0, spp, seed)
    for block, world in ((8, 3), (4, 2), (32, 5)):
        tiles = e.new_accum(W, H)
        for rank in range(world):
            y0, y1, stride = interleaved_window(H, rank, world, block)
            e.set_row_interleave(block, stride)
            fn(tiles, 0, spp, seed, y0=y0, y1=y1)
        e.set_row_interleave(0)
        assert torch.equal(tiles, full), (block, world)
    with pytest.raises(Exception, match="power of two"):
        e.set_row_interleave(6, 12)


def test_progressive_preview_never_stalls_and_shows_finished_frames(monkeypatch):
    """The output step after the hot path: render(enable_preview=True) tonemaps on the device and
    copies the 8-bit frame to pinned memory on a side stream (renderer._preview_frame_async); the window -- a stub
    here, there is no display on the box -- receives only completed frames, each equal to buffer_to_image of the
    accumulation buffer after some whole number of samples, in increasing order, and the final image is unaffected."""
    from path_tracer_python_b200 import preview

    shown = []

    class StubPreview:
        def __init__(self, w, h, title=""):
            self.size = (w, h)

        def start(self, update_interval_ms=50):
            pass

        def update_image(self, img):
            shown.append(np.array(img, copy=True))

        def process_events(self):
            torch.cuda.synchronize()          # "the GUI took a while": lets the side-stream copy finish

        def finish(self):
            pass

    monkeypatch.setattr(preview, "LivePreview", StubPreview)
    _, gmeta, out = golden_fixtures.load_fixture("wavefront_comparison")
    spp = 6
    r, cs, meta = _renderer(spp=spp, seed=gmeta["rng_seed"])
    r.background_color = gmeta["background"]
    r.render(enable_preview=True)                      # one sample per launch, a preview frame after each
    final = r.accum_numpy()
    # reference frames: the accumulation buffer after k samples, tonemapped like preview.py:117-132
    e = r.engine
    frames = []
    acc = e.new_accum(r.cam.img_width, r.cam.img_height)
    for k in range(spp):
        e.render_megakernel(acc, k, 1, gmeta["rng_seed"])
        frames.append(preview.buffer_to_image(acc.cpu().numpy(), k + 1))
    assert np.array_equal(final, acc.cpu().numpy())
    assert len(r.sample_times) == spp and 1 <= len(shown) <= spp
    last = -1
    for img in shown:
        match = [k for k, f in enumerate(frames) if np.array_equal(f, img)]
        assert match, "a preview frame that is not the tonemap of any whole-sample state"
        assert match[0] > last - 1
        last = match[0]
    # the async path itself: nothing ready on the first call, the first frame on a later one
    r2, _, _ = _renderer(spp=2, seed=1)
    r2.engine.render_megakernel(r2._accum, 0, 1, 1)
    r2.current_sample = 1
    want = preview.buffer_to_image(r2.accum_numpy(), 1)
    assert r2._preview_frame_async() is None
    torch.cuda.synchronize()
    r2.engine.render_megakernel(r2._accum, 1, 1, 1)
    r2.current_sample = 2
    got = r2._preview_frame_async()
    assert got is not None and np.array_equal(got, want)
