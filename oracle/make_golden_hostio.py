"""TEST INFRASTRUCTURE ONLY — golden vectors for the host-side helpers of the entry points, recorded from the
untouched reference ``utils/utils.py`` (img_resize :90-101, load_segment :104-153) in the build container:

    python -m oracle.make_golden_hostio      ->  tests/golden/hostio.npz
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("_vst_reference_utils", os.path.join(ref_shim.REF_ROOT, "utils", "utils.py"))
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)
    rng = np.random.default_rng(5)
    table = np.array([(0, 0, 255), (0, 255, 0), (0, 0, 0), (255, 255, 255), (255, 0, 0), (255, 255, 0), (128, 128, 128),
                      (0, 255, 255), (255, 0, 255)], np.uint8)
    # exact table colours, slightly perturbed ones (nearest-colour rule) and arbitrary colours
    seg = table[rng.integers(0, 9, (24, 20))].astype(np.int32)
    seg[8:16] += rng.integers(-20, 21, (8, 20, 3))
    seg[16:] = rng.integers(0, 256, (8, 20, 3))
    seg = np.clip(seg, 0, 255).astype(np.uint8)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "seg.png")
        Image.fromarray(seg).save(path)
        labels = ru.load_segment(path)
        labels_resized = ru.load_segment(path, size=(10, 12))
    img = Image.fromarray(rng.integers(0, 256, (37, 53, 3), dtype=np.uint8))
    r1 = np.array(ru.img_resize(img, 40, down_scale=4))
    r2 = np.array(ru.img_resize(img, 1280, down_scale=4))
    np.savez_compressed(os.path.join(OUT, "hostio.npz"), seg=seg, labels=labels, labels_resized=labels_resized,
                        img=np.array(img), resized_40=r1, resized_1280=r2)
    print("wrote hostio.npz", labels.shape, labels_resized.shape, r1.shape, r2.shape)


if __name__ == "__main__":
    main()
