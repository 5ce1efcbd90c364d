#!/usr/bin/env python
"""Write configs/*.json: the frozen synthetic inputs of SURVEY 8d (C1..C5b of tests/scenes.py) in a language-neutral form --
the generating parameters AND the three uniform blocks they produce (floats as written by mm_host_sky / mm_host_camera, which are
byte-identical to the reference's SkyManager / Camera), so that any harness can bind exactly the same inputs."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg, scenes

def main():
    mm = _pkg.load_package()
    assets = scenes.load_assets()
    out = os.path.join(ROOT, "configs")
    os.makedirs(out, exist_ok=True)
    for name, cfg in scenes.CONFIGS.items():
        sc = scenes.make_scene(mm, name, assets)
        doc = {
            "name": name, "width": sc["W"], "height": sc["H"],
            "camera": {"position": list(map(float, cfg["pos"])), "yaw": float(cfg["yaw"]), "pitch": float(cfg["pitch"]), "fov_deg": 45.0, "aspect": 1920.0 / 1080.0},
            "sun": {"elevation": float(cfg["elevation"]), "azimuth": float(cfg["azimuth"])},
            "sky": {"wind": list(map(float, cfg["wind"])), "time": float(cfg["time"]), "turbidity": 10.0, "rayleigh": 2.0, "mie": 0.005, "mie_directional": 0.8},
            "placement": "shipped CloudPlacement.png" if cfg["placement"] == "shipped" else {"constant_R_coverage": cfg["placement"][0], "constant_B_type": cfg["placement"][1]},
            "volumes": "shipped lowResCloudShape 128^3, hiResCloudShape 32^3, CurlNoiseFBM 128^2 (tests/golden/assets)",
            "uniform_blocks_f32": {"UniformCameraObject_160B": [float(x) for x in sc["cam"]], "UniformSunObject_116B": [float(x) for x in sc["sun"]],
                                   "UniformSkyObject_52B": [float(x) for x in sc["sky"]]},
        }
        with open(os.path.join(out, name + ".json"), "w") as f:
            json.dump(doc, f, indent=1)
        print("wrote", name)

if __name__ == "__main__":
    main()
