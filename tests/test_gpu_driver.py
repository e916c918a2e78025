"""The C++ driver (tiny-path-tracer_b200/host/main.cpp = the reference's main() flow with the sample
loop replaced): reads ./config.ini with the reference's keys, renders on the GPU, writes img.ppm
(+ bonus pictures) in the reference's P3 formats. Compared with the reference executable's own
output for the same config: both are Monte-Carlo renders of the same scene, so the 8-bit pictures
agree up to noise."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIG = """[DEFAULT]
width = 120
height = 120
sample = 256
recur_depth = 15
fov = 90.0
allow_bonus_pic = 1
bonus_pic = 4
[BLUR]
aperture = 0.1
[CAM_MOTION]
start_time = 0.0
end_time = 0.0
"""


def read_p3(path):
    tok = open(path).read().split()
    assert tok[0] == "P3"
    w, h = int(tok[1]), int(tok[2])
    return np.array(tok[4:], np.int32).reshape(h, w, 3)


def test_driver_writes_reference_compatible_pictures(T, gpu, tmp_path):
    exe = os.path.join(T.LIB_DIR, "Path_tracer_b200")
    assert os.path.exists(exe), "driver not built (make -C tiny-path-tracer_b200/host driver)"
    (tmp_path / "config.ini").write_text(CONFIG)
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "time: " in out.stdout and "[DEFAULT]" in out.stdout  # ini echo + the reference's time line
    img = read_p3(tmp_path / "img.ppm")
    assert img.shape == (120, 120, 3) and img.min() >= 0 and img.max() <= 255
    # main picture: one pixel per line; bonus pictures: one long line (main.cpp:183-189 vs :210-211)
    assert len(open(tmp_path / "img.ppm").read().splitlines()) == 3 + 120 * 120
    for k in range(4):
        assert len(open(tmp_path / f"img_{k}.ppm").read().splitlines()) == 4
    last = read_p3(tmp_path / "img_3.ppm")
    assert np.array_equal(last, img)  # the last bonus picture uses all samples
    ref_exe = os.path.join(T.REPO_ROOT, "oracle", "_ref", "Path_tracer")
    if os.path.exists(ref_exe):  # the unmodified reference executable, same config.ini
        ref_dir = tmp_path / "ref"
        ref_dir.mkdir()
        shutil.copy(tmp_path / "config.ini", ref_dir / "config.ini")
        r = subprocess.run([ref_exe], cwd=ref_dir, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0
        ref = read_p3(ref_dir / "img.ppm")
        rmse = np.sqrt(np.mean((img.astype(np.float64) - ref) ** 2)) / 255.0
        assert rmse < 0.06, rmse  # 256 spp on both sides; the reference's row-correlated noise dominates
        assert abs(img.mean() - ref.mean()) < 0.03 * ref.mean()
