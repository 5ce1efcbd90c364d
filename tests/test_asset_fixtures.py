"""The committed texture fixtures (tests/golden/assets, made by tools/make_asset_fixtures.py) against the reference's OWN decoder on the
reference's OWN files: Libraries/stb/stb_image.h compiled into oracle/_ref/libref_stb.so and called as Texture::initFromFile /
Texture3D::initFromFile call it (Texture.cpp:212-246, 502-538: STBI_rgb_alpha; slice i -> z = i).  Runs where /root/reference exists."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_STB = os.path.join(ROOT, "oracle", "_ref", "libref_stb.so")
TEX = "/root/reference/SkyEngine/SkyEngine/Textures/"


def test_fixture_hashes_match_the_manifest(assets):
    for key, name in (("placement", "CloudPlacement"), ("curl", "CurlNoiseFBM"), ("lowres", "lowResCloudShape"), ("hires", "hiResCloudShape")):
        assert hashlib.sha256(assets[key].tobytes()).hexdigest() == assets["manifest"][name]["sha256"]
    assert assets["lowres"].shape == (128, 128, 128, 4) and assets["hires"].shape == (32, 32, 32, 4)
    assert (assets["hires"][..., 3] == 0).all()                       # the hi-res volume's alpha is all 0


@pytest.mark.skipif(not (os.path.exists(REF_STB) and os.path.isdir(TEX)), reason="needs /root/reference and oracle/_ref/libref_stb.so")
def test_fixtures_equal_the_reference_decoder_on_the_reference_files(assets):
    lib = C.CDLL(REF_STB)
    lib.ref_stbi_load_rgba.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_size_t]

    def load(path, shape):
        out, w, h = np.zeros(shape, np.uint8), C.c_int(), C.c_int()
        assert lib.ref_stbi_load_rgba(path.encode(), C.byref(w), C.byref(h), out.ctypes.data, out.nbytes) == 0, path
        assert (h.value, w.value) == shape[:2]
        return out
    assert np.array_equal(load(TEX + "CloudPlacement.png", (512, 512, 4)), assets["placement"])            # VulkanApplication.cpp:254
    assert np.array_equal(load(TEX + "CurlNoiseFBM.png", (128, 128, 4)), assets["curl"])                   # :258 (the app loads the PNG)
    assert np.array_equal(load(TEX + "CurlNoiseFBM.tga", (128, 128, 4)), assets["curl"])                   # the generator's own output
    for i in range(128):                                                                                   # :260, Texture.cpp:509-523
        assert np.array_equal(load(TEX + f"3DTextures/lowResCloudShape/lowResCloud({i}).tga", (128, 128, 4)), assets["lowres"][i]), i
    for i in range(32):                                                                                    # :262
        assert np.array_equal(load(TEX + f"3DTextures/hiResCloudShape/hiResClouds ({i}).tga", (32, 32, 4)), assets["hires"][i]), i
    import scenes
    assert np.array_equal(load(TEX + "NightSky/nightSky_noOrange.png", (1080, 1920, 4)), scenes.shipped_night_sky())   # :255-256
