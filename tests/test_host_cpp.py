"""The C++ host side (monte_carlo_localization_b200/host): map + config loaders against their
Python twins on CPU; the ParticleFilter mirror / replay driver on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from monte_carlo_localization_b200 import maps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "monte_carlo_localization_b200", "host")
REF = "/root/reference"


def _host():
    so = os.path.join(HOST, "libpf_host.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", HOST])
    L = C.CDLL(so)
    L.pfhost_load_map.argtypes = [C.c_char_p, C.POINTER(C.c_int8), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_char_p, C.c_int]
    L.pfhost_load_config.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
    return L


def _load_cpp(path):
    L = _host()
    w, h, res = C.c_int(), C.c_int(), C.c_float()
    org = (C.c_double * 3)()
    err = C.create_string_buffer(256)
    rc = L.pfhost_load_map(path.encode(), None, 0, C.byref(w), C.byref(h), C.byref(res), org, err, 256)
    assert rc == 0, err.value
    data = np.empty(w.value * h.value, dtype=np.int8)
    rc = L.pfhost_load_map(path.encode(), data.ctypes.data_as(C.POINTER(C.c_int8)), data.size, C.byref(w), C.byref(h),
                           C.byref(res), org, err, 256)
    assert rc == 0
    return data.reshape(h.value, w.value), np.float32(res.value), tuple(org)


def _write_yaml(path, image, flow=True, negate=0, occ=0.65, free=0.196, res=0.05, origin=(-1.5, 2.25, 0.0)):
    with open(path, "w") as f:
        f.write("image: %s\nresolution: %r\n" % (image, res))
        if flow:
            f.write("origin: [%r, %r, %r]\n" % origin)
        else:
            f.write("origin:\n- %r\n- %r\n- %r\n" % origin)
        f.write("negate: %d\noccupied_thresh: %r\nfree_thresh: %r\n" % (negate, occ, free))


@pytest.mark.parametrize("mode,ext", [("L", "png"), ("RGB", "png"), ("RGBA", "png"), ("LA", "png"), ("P", "png"),
                                     ("L", "pgm")])
@pytest.mark.parametrize("flow", [True, False])
def test_cpp_loader_matches_python_loader_on_generated_images(tmp_path, mode, ext, flow):
    from PIL import Image
    rng = np.random.default_rng(7)
    h, w = 37, 53
    gray = rng.choice(np.array([0, 30, 100, 205, 230, 254, 255], dtype=np.uint8), size=(h, w))
    if mode == "L":
        im = Image.fromarray(gray, "L")
    elif mode == "RGB":
        im = Image.fromarray(np.stack([gray, np.roll(gray, 1, 0), np.roll(gray, 1, 1)], -1), "RGB")
    elif mode == "RGBA":
        a = rng.choice(np.array([0, 128, 255], dtype=np.uint8), size=(h, w))
        im = Image.fromarray(np.stack([gray, gray, np.roll(gray, 2, 1), a], -1), "RGBA")
    elif mode == "LA":
        a = rng.choice(np.array([0, 255], dtype=np.uint8), size=(h, w))
        im = Image.fromarray(np.stack([gray, a], -1), "LA")
    else:
        im = Image.fromarray(gray, "L").convert("P")
    img = "m.%s" % ext
    im.save(str(tmp_path / img))
    y = str(tmp_path / "m.yaml")
    _write_yaml(y, img, flow=flow, negate=int(flow))
    data, res, org = _load_cpp(y)
    g = maps.load_map_yaml(y)
    assert np.array_equal(data, g.data)
    assert res == g.resolution and org == g.origin


@pytest.mark.skipif(not os.path.exists(REF + "/maps/sibal1.yaml"), reason="reference maps absent")
@pytest.mark.parametrize("yml", ["sibal1.yaml", "Spielberg_map.yaml", "basement_fixed.map.yaml", "first_map.yaml",
                                 "new_map1.yaml", "redbull_1.yaml", "icra_2_clean.yaml", "slam_map.yaml",
                                 "map_1753950572.yaml", "map_1755669035.yaml"])
def test_cpp_loader_on_every_shipped_map(yml):
    path = os.path.join(REF, "maps", yml)
    data, res, org = _load_cpp(path)
    g = maps.load_map_yaml(path)
    assert np.array_equal(data, g.data)
    assert res == g.resolution and org == g.origin


def test_cpp_loader_reports_errors(tmp_path):
    L = _host()
    w, h, res = C.c_int(), C.c_int(), C.c_float()
    org = (C.c_double * 3)()
    err = C.create_string_buffer(256)
    assert L.pfhost_load_map(str(tmp_path / "missing.yaml").encode(), None, 0, C.byref(w), C.byref(h), C.byref(res),
                             org, err, 256) == -1
    assert b"cannot open" in err.value
    y = str(tmp_path / "m.yaml")
    _write_yaml(y, "absent.png")
    assert L.pfhost_load_map(y.encode(), None, 0, C.byref(w), C.byref(h), C.byref(res), org, err, 256) == -1


def test_config_parser_reads_reference_surface(tmp_path):
    """The shipped mcl_config.yaml values (config/mcl_config.yaml:3-58); vestigial keys ignored."""
    path = os.path.join(REF, "config", "mcl_config.yaml")
    if not os.path.exists(path):
        path = str(tmp_path / "mcl_config.yaml")
        with open(path, "w") as f:
            f.write("particle_filter:\n  ros__parameters:\n    max_particles: 2000  # n\n    max_viz_particles: 60\n"
                    "    max_range: 12.0\n    delay_compensation_factor: 3.5\n    sim_mode: false\n"
                    "    motion_dispersion_x: 0.05\n    motion_dispersion_y: 0.025\n    motion_dispersion_theta: 0.25\n"
                    "    lidar_offset_x: 0.288\n    z_hit: 0.80\n    z_short: 0.01\n    z_max: 0.07\n    z_rand: 0.12\n"
                    "    sigma_hit: 8.0\n    range_method: \"cddt\"\n    angle_step: 18\n    squash_factor: 2.2\n"
                    "    num_threads: 3\n    timer_frequency: 200.0\nmap_server:\n  ros__parameters:\n    map: 'sibal1'\n")
    out = (C.c_double * 17)()
    assert _host().pfhost_load_config(path.encode(), out) == 0
    got = list(out)
    want = [2000, 18, 60, 2.2, 12.0, 0.01, 0.07, 0.12, 0.80, 8.0, 0.05, 0.025, 0.25, 0.288, 200.0, 3, 3.5]
    assert got == want


@pytest.mark.gpu
def test_cpp_particle_filter_replay_tracks(tmp_path):
    """mcl_replay: C++ ParticleFilter (lidarCB -> update -> MCL -> expected_pose) from a map yaml."""
    from PIL import Image
    g = maps.load_named_map("sibal1")
    img = np.where(g.data[::-1] == 100, 0, np.where(g.data[::-1] == 0, 254, 205)).astype(np.uint8)
    Image.fromarray(img, "L").save(str(tmp_path / "sibal1.png"))
    y = str(tmp_path / "sibal1.yaml")
    _write_yaml(y, "sibal1.png", flow=False, occ=0.65, free=0.1, res=float(g.resolution), origin=g.origin)
    exe = os.path.join(HOST, "mcl_replay")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", HOST])
    res = subprocess.run([exe, y, "--particles", "4000", "--steps", "15", "--x", "-3.3", "--y", "1.6", "--theta", "0.3"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert res.returncode == 0, res.stdout
    assert "MAX_RANGE_PX 239" in res.stdout and "iter  15" in res.stdout
    # the estimate must TRACK the ground truth, not merely exist: every tick's position error, as printed
    import re
    errs = [float(m) for m in re.findall(r"err ([0-9.]+) m", res.stdout)]
    assert len(errs) == 15 and max(errs[3:]) < 0.3, errs
