"""The OpenEXR reader / writer for baked shading maps (iris_b200/utils/exr.py): round trips, and both directions against OpenCV --
the library the reference uses for these files (bake_shading.py:131, utils/dataset/*) -- when this OpenCV build has EXR support."""
import os

import numpy as np
import pytest

from iris_b200.utils import exr


@pytest.mark.parametrize("compression", ["none", "zips", "zip"])
@pytest.mark.parametrize("half", [False, True])
def test_exr_round_trip(tmp_path, compression, half):
    rng = np.random.default_rng(1)
    img = (rng.random((37, 53, 3)) * 20.0).astype(np.float32)      # HDR values, a height that is not a multiple of the 16-line ZIP block
    img[3, 4] = [0.0, 1e-8, 6.0e4]
    p = str(tmp_path / "a.exr")
    exr.write_exr(p, img, half=half, compression=compression)
    back = exr.read_exr(p)
    want = img.astype(np.float16).astype(np.float32) if half else img
    assert back.dtype == np.float32 and np.array_equal(back, want)
    assert np.array_equal(exr.read_exr(p, "B")[..., 0], want[..., 2])
    # a smooth image must actually shrink under ZIP (the predictor is applied, not just deflate)
    if compression == "zip":
        smooth = np.tile(np.linspace(0, 1, 53, dtype=np.float32)[None, :, None], (37, 1, 3))
        exr.write_exr(p, smooth, compression="zip")
        assert os.path.getsize(p) < 0.5 * smooth.nbytes and np.array_equal(exr.read_exr(p), smooth)


def test_exr_single_channel_and_errors(tmp_path):
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    p = str(tmp_path / "y.exr")
    exr.write_exr(p, a)
    assert np.array_equal(exr.read_exr(p)[..., 0], a)
    open(p, "wb").write(b"not an exr file at all")
    with pytest.raises(ValueError):
        exr.read_exr(p)


def test_exr_against_opencv(tmp_path):
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    img = (rng.random((48, 40, 3)) * 5.0).astype(np.float32)
    p = str(tmp_path / "cv.exr")
    try:
        ok = cv2.imwrite(p, img[:, :, [2, 1, 0]])                  # the reference's call (BGR in)
    except cv2.error:
        ok = False
    if not ok:
        pytest.skip("this OpenCV build has no OpenEXR codec")
    assert np.array_equal(exr.read_exr(p), img)                    # OpenCV's ZIP file -> our reader, RGB out
    exr.write_exr(p, img)
    assert np.array_equal(cv2.imread(p, -1)[:, :, [2, 1, 0]], img)   # our file -> the reference's reader
