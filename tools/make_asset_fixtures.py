#!/usr/bin/env python
"""Decode the reference's texture assets into the byte layout the engine uploads.

Test/bench infrastructure only.  Runs in the build container (reads /root/reference,
which does not exist on the GPU box) and writes tests/golden/assets/*.npz, which are
committed so that parity tests, smoke() and bench.py can bind "the same textures" the
reference binds (VulkanApplication.cpp:253-262).

Decode order follows Texture.cpp:212-246 (2D: stbi_load(..., STBI_rgb_alpha), top-down
rows, x fastest) and Texture.cpp:502-538 (3D: slice i -> z = i, z-major).  PIL's decode
of these five assets is byte-identical to stb_image 2.16.
"""
import hashlib, json, os, sys
import numpy as np
from PIL import Image

REF = "/root/reference/SkyEngine/SkyEngine/Textures/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "assets")

def rgba(path):
    return np.array(Image.open(path).convert("RGBA"), dtype=np.uint8)

def main():
    os.makedirs(OUT, exist_ok=True)
    low = np.stack([rgba(REF + f"3DTextures/lowResCloudShape/lowResCloud({i}).tga") for i in range(128)])
    hi = np.stack([rgba(REF + f"3DTextures/hiResCloudShape/hiResClouds ({i}).tga") for i in range(32)])
    placement = rgba(REF + "CloudPlacement.png")
    curl = rgba(REF + "CurlNoiseFBM.tga")
    curl_png = rgba(REF + "CurlNoiseFBM.png")
    assert (curl[..., :3] == curl_png[..., :3]).all()
    night = rgba(REF + "NightSky/nightSky_noOrange.png")          # the star map the app binds (VulkanApplication.cpp:255-256): 1920x1080, not a power of two
    man = {}
    for name, arr in [("lowResCloudShape", low), ("hiResCloudShape", hi),
                      ("CloudPlacement", placement), ("CurlNoiseFBM", curl), ("nightSky_noOrange", night)]:
        np.savez_compressed(os.path.join(OUT, name + ".npz"), rgba8=arr)
        man[name] = {"shape": list(arr.shape), "sha256": hashlib.sha256(arr.tobytes()).hexdigest()}
        print(name, arr.shape, man[name]["sha256"][:16])
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1)

if __name__ == "__main__":
    sys.exit(main())
