"""CPU tests of the boundary: the C-ABI library loads and exports every declared symbol,
fails loudly without a device, and the map loaders agree."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from monte_carlo_localization_b200 import capi, maps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mcl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mcl_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), "libmcl_b200.so does not export %s" % name
    # the ctypes table binds exactly the header's functions
    assert sorted(capi.SIGNATURES) == declared
    assert L.mcl_abi_version() == 3


def test_params_defaults_match_reference_declarations():
    """src/particle_filter.cpp:23-37 defaults."""
    p = capi.default_params()
    assert (p.max_particles, p.max_viz_particles, p.angle_step) == (2000, 60, 18)
    assert (p.squash_factor, p.max_range) == (2.2, 12.0)
    assert (p.z_short, p.z_max, p.z_rand, p.z_hit, p.sigma_hit) == (0.01, 0.07, 0.12, 0.80, 8.0)
    assert (p.motion_dispersion_x, p.motion_dispersion_y, p.motion_dispersion_theta) == (0.05, 0.025, 0.25)
    assert p.num_filters == 1


def test_no_cpu_fallback_without_device():
    L = capi.load_library()
    if L.mcl_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(capi.MclError) as e:
        capi.MclContext(max_particles=100)
    assert e.value.status == capi.MCL_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_invalid_arguments_are_rejected_before_any_device_work():
    L = capi.load_library()
    h = C.c_void_p()
    p = capi.default_params(max_particles=0)
    assert L.mcl_create(C.byref(p), 0, C.byref(h)) == capi.MCL_ERR_INVALID
    assert L.mcl_create(None, 0, C.byref(h)) == capi.MCL_ERR_INVALID
    assert b"invalid" in L.mcl_status_str(capi.MCL_ERR_INVALID)


def test_product_never_touches_the_oracle():
    """The product package must not import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "monte_carlo_localization_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src.replace("the oracle's in CPU tests", "").replace("CPU oracle", ""), \
                    "%s mentions oracle" % os.path.join(dp, f)


def test_map_fixtures_match_survey_counts():
    want = {"sibal1": (350, 177, 239, (26948, 35002, 0)),
            "Spielberg_map": (2000, 2000, 207, (3960078, 33998, 5924)),
            "basement_fixed": (1300, 1300, 238, (275742, 14374, 1399884))}
    for name, (w, h, m, counts) in want.items():
        g = maps.load_named_map(name)
        assert (g.width, g.height, g.max_range_px(), g.counts()) == (w, h, m, counts)


@pytest.mark.skipif(not os.path.exists("/root/reference/maps/sibal1.yaml"), reason="reference maps absent")
def test_yaml_loader_reproduces_fixtures():
    for name, yml in (("sibal1", "sibal1.yaml"), ("Spielberg_map", "Spielberg_map.yaml"), ("first_map", "first_map.yaml")):
        g = maps.load_map_yaml(os.path.join("/root/reference/maps", yml))
        f = maps.load_named_map(name)
        assert np.array_equal(g.data, f.data) and g.origin == f.origin and g.resolution == f.resolution


def test_trinary_rule():
    img = np.array([[0, 255, 205, 128]], dtype=np.uint8)
    g = maps.image_to_grid(img, negate=False, occupied_thresh=0.65, free_thresh=0.196)
    assert g.tolist() == [[100, 0, -1, -1]]
    g = maps.image_to_grid(img, negate=True, occupied_thresh=0.65, free_thresh=0.196)
    assert g.tolist() == [[0, 100, 100, -1]]
    # rows flip: grid row 0 is the image's bottom row
    img2 = np.array([[0], [255]], dtype=np.uint8)
    assert maps.image_to_grid(img2, False, 0.65, 0.196).ravel().tolist() == [0, 100]


def test_synthetic_levine_stand_in():
    """The procedural stand-in for the missing levine.pgm honours maps/levine.yaml's geometry."""
    from monte_carlo_localization_b200 import synth
    g = maps.synth_levine()
    assert (g.width, g.height) == (2048, 2048) and g.max_range_px() == 239
    assert g.origin[:2] == (-51.224998, -51.224998)
    free, occ, unk = g.counts()
    assert free > 100000 and occ > 5000 and unk > 3000000
    x, y, th = synth.centreline_loop(g, n_points=200, mask=(g.data == 0))
    rows = ((y - g.origin[1]) / g.resolution_f64).astype(int)
    cols = ((x - g.origin[0]) / g.resolution_f64).astype(int)
    assert (g.data[rows, cols] == 0).mean() > 0.95      # the ground-truth loop runs through free space
