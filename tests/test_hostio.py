"""CPU: host-side helpers of the entry points against vectors recorded from the reference's utils/utils.py
(oracle/make_golden_hostio.py): load_segment's colour -> label rule and img_resize."""
import os

import numpy as np
from PIL import Image

from tests.helpers import load_golden
from vstnet_b200.hostio import img_resize, labels_from_colors, load_segment


def test_labels_from_colors_matches_reference_load_segment(tmp_path):
    g = load_golden("hostio.npz")
    # ties between equidistant table colours are resolved by the reference from a neighbouring pixel's label
    # (utils.py:128-132, itself guarded by a bare except); only unambiguous pixels are compared
    seg = g["seg"].astype(np.int32)
    table = np.array([(0, 0, 0), (255, 255, 255), (0, 255, 0), (0, 0, 255), (255, 0, 0), (255, 255, 0), (128, 128, 128),
                      (0, 255, 255), (255, 0, 255)], np.int32)
    d = np.sort(np.abs(seg[:, :, None, :] - table[None, None]).sum(-1), axis=-1)
    unambiguous = d[..., 0] < d[..., 1]
    assert unambiguous.mean() > 0.9
    out = labels_from_colors(g["seg"])
    assert out.dtype == np.uint8 and out.shape == g["labels"].shape
    assert np.array_equal(out[unambiguous], g["labels"][unambiguous])
    path = os.path.join(str(tmp_path), "seg.png")
    Image.fromarray(g["seg"]).save(path)
    assert np.array_equal(load_segment(path)[unambiguous], g["labels"][unambiguous])
    assert load_segment(path, size=(10, 12)).shape == g["labels_resized"].shape
    assert load_segment(os.path.join(str(tmp_path), "missing.png")) is None


def test_img_resize_matches_reference():
    g = load_golden("hostio.npz")
    img = Image.fromarray(g["img"])
    assert np.array_equal(np.array(img_resize(img, 40, down_scale=4)), g["resized_40"])
    assert np.array_equal(np.array(img_resize(img, 1280, down_scale=4)), g["resized_1280"])


def test_oracle_seg_remapping_matches_reference():
    """oracle.seg_self_remapping / seg_cross_remapping against vectors recorded from the reference's SegReMapping
    (oracle/make_golden_seg.py, synthetic relation table)."""
    from oracle import vst_oracle as O
    g = load_golden("seg_remap.npz")
    table = g["table"]
    for k in range(3):
        r = float(g["ratio%d" % k])
        cs = O.seg_self_remapping(g["c%d" % k], table, r)
        ss = O.seg_self_remapping(g["s%d" % k], table, r)
        assert np.array_equal(cs, g["c_self%d" % k]) and np.array_equal(ss, g["s_self%d" % k])
        assert np.array_equal(O.seg_cross_remapping(cs, ss, table), g["c_cross%d" % k])
        assert not np.array_equal(cs, g["c%d" % k]) or k == 2            # the small labels really moved
    h = load_golden("hostio.npz")
    seg = h["seg"].astype(np.int32)
    tab = np.array([c for c, _ in O.SEG_COLOR_TABLE], np.int32)
    d = np.sort(np.abs(seg[:, :, None, :] - tab[None, None]).sum(-1), axis=-1)
    unambiguous = d[..., 0] < d[..., 1]
    assert np.array_equal(O.seg_labels_from_colors(h["seg"])[unambiguous], h["labels"][unambiguous])
    assert np.array_equal(O.seg_labels_from_colors(h["seg"]), labels_from_colors(h["seg"]))
