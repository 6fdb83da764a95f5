"""oracle/cv2_chain.py restates the reference's glue code; prove it equals the reference's own
classes whenever /root/reference is mounted (this container; not the GPU box)."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "preprocess")), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref_mod():
    pytest.importorskip("cv2")
    sys.dont_write_bytecode = True
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    sys.path.insert(0, REF)
    try:
        import src.preprocess as sp
        import src.preprocess.registry as reg
        yield sp, reg
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_cv2_chain_equals_reference_classes(ref_mod):
    from oracle import cv2_chain as R
    sp, reg = ref_mod
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, (144, 256, 3)).astype(np.uint8)
    for params, k in [({"space": "LAB", "clip_limit": 2.0, "tile_grid": 8}, 3),
                      ({"space": "ycrcb", "clip_limit": "3", "tile_grid": 1}, 6),
                      ({"space": "hsv", "clip_limit": 0, "tile_grid": 8.9}, 11),
                      ({}, 1)]:
        cfg = {"chain": [{"name": "CLAHEDehaze", "params": params}, {"name": "MedianDerain", "params": {"ksize": k}}]}
        want = sp.PreprocessPipeline(cfg)(img)
        space, clip, grid = R.coerce_clahe_params(params)
        got = R.chain(img, space, clip, grid, R.coerce_ksize({"ksize": k}))
        assert np.array_equal(got, want), (params, k)


def test_registry_names_match_reference(ref_mod):
    import rvb200
    _, reg = ref_mod
    assert set(reg.REGISTRY) <= set(rvb200.REGISTRY)
    with pytest.raises(KeyError) as e1:
        reg.get_op_class("Nope")
    with pytest.raises(KeyError) as e2:
        rvb200.get_op_class("Nope")
    assert "Preprocess op 'Nope' not found" in str(e1.value) and "Preprocess op 'Nope' not found" in str(e2.value)


def test_param_coercions_match_reference_source():
    """Our coerce() helpers against the coercions spelled out in clahe_dehaze.py:14-17 / median_derain.py:11-13."""
    from rvb200.preprocess.ops import clahe_dehaze as c, median_derain as m
    from oracle import cv2_chain as R
    for params in [{}, {"space": "Lab"}, {"space": "hsv"}, {"tile_grid": 1}, {"tile_grid": "8"}, {"tile_grid": 8.9},
                   {"clip_limit": "2"}, {"clip_limit": 0}, {"clip_limit": -1}, {"bogus": 1}]:
        space, clip, grid = R.coerce_clahe_params(params)
        assert c.coerce(params) == ("LAB" if space == "LAB" else "YCrCb", clip, grid)
    for k in [1, 2, 3, 3.9, "5", 6, 7, 8, 9, 11, 100]:
        assert m.coerce({"ksize": k}) == R.coerce_ksize({"ksize": k})
