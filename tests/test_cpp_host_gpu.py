"""The C++ host mirror (host/ComputeShader.h, SkyManager, Camera) drives the same C-ABI: the headless frame_demo
binary must produce the frame the Python mirror produces, bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import scenes


def test_cpp_demo_fails_loudly_without_gpu(mm, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    demo = os.path.join(os.path.dirname(mm.library_path()), "frame_demo")
    r = subprocess.run([demo, "a", "b", "c", "d", "8", "8", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode != 0


@pytest.mark.gpu
def test_cpp_demo_matches_python_mirror(mm, assets, tmp_path):
    for k in ("placement", "curl", "lowres", "hires"):
        assets[k].tofile(tmp_path / f"{k}.raw")
    W, H = 192, 108
    demo = os.path.join(os.path.dirname(mm.library_path()), "frame_demo")
    out = tmp_path / "out.f32"
    r = subprocess.run([demo] + [str(tmp_path / f"{k}.raw") for k in ("placement", "curl", "lowres", "hires")] + [str(W), str(H), str(out), "0.25", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out, np.float32).reshape(H, W, 4)
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    cs.setFilterMode(mm.MM_FILTER_HYBRID)
    want = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    cs.close()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
