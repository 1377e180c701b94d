"""TEST INFRASTRUCTURE ONLY — golden vectors for the mask-preparation path, recorded from the untouched reference
``models/segmentation/SegReMapping.py`` (numpy class, the one image_transfer.py / video_transfer.py call) in the build
container, on a SYNTHETIC relation table (the algorithm is table-agnostic; the reference's ade20k table is not copied):

    python -m oracle.make_golden_seg      ->  tests/golden/seg_remap.npz
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synthetic_table(rng, n_classes=150):
    """Like the reference's table: column c lists all classes from the most to the least related, c itself last."""
    t = np.zeros((n_classes, n_classes), np.int64)
    for c in range(n_classes):
        others = [k for k in rng.permutation(n_classes) if k != c]
        t[:, c] = others + [c]
    return t


def label_map(rng, h, w, big, small):
    """Blocky map of the `big` labels with a few tiny patches of the `small` ones."""
    seg = np.zeros((h, w), np.uint8)
    ys, xs = np.linspace(0, h, 4).astype(int), np.linspace(0, w, len(big) // 3 + 2).astype(int)
    k = 0
    for i in range(3):
        for j in range(len(xs) - 1):
            seg[ys[i]:ys[i + 1], xs[j]:xs[j + 1]] = big[k % len(big)]
            k += 1
    for l in small:
        y, x = rng.integers(0, h - 3), rng.integers(0, w - 3)
        seg[y:y + 2, x:x + 3] = l
    return seg


def main():
    spec = importlib.util.spec_from_file_location("_vst_reference_segremap",
                                                  os.path.join(ref_shim.REF_ROOT, "models", "segmentation", "SegReMapping.py"))
    rs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rs)
    rng = np.random.default_rng(11)
    table = synthetic_table(rng)
    out = {"table": table.astype(np.int32)}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "rel.npy")
        np.save(path, table)
        for case, (h, w, big_c, small_c, big_s, small_s, ratio) in enumerate([
                (96, 128, [2, 4, 9, 16, 21], [7, 33, 140], [2, 9, 16, 50, 77], [4], 0.01),
                (60, 44, [0, 1, 149], [5, 6, 7, 8], [1, 3, 100], [], 0.02),
                (128, 160, [10, 20, 30, 40, 50, 60], [], [10, 20, 31, 41], [99], 0.05)]):
            rm = rs.SegReMapping(path, min_ratio=ratio)
            c = label_map(rng, h, w, big_c, small_c)
            s = label_map(rng, h + 8, w - 4, big_s, small_s)
            c_self = rm.self_remapping(c)
            s_self = rm.self_remapping(s)
            c_cross = rm.cross_remapping(c_self, s_self)
            out.update({"c%d" % case: c, "s%d" % case: s, "c_self%d" % case: c_self, "s_self%d" % case: s_self,
                        "c_cross%d" % case: c_cross, "ratio%d" % case: np.float32(ratio)})
    np.savez_compressed(os.path.join(OUT, "seg_remap.npz"), **out)
    print("wrote seg_remap.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
