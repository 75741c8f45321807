#!/usr/bin/env python
"""Render a small view with the CPU oracle (ReSTIR DI, accumulated) and write a tone-mapped PNG: a visual check of the environment
light and the image textures that needs no GPU.  Usage: tools/render_oracle_png.py out.png [width height frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gfxexp_b200 import abi, engine, imageio, scenes
from tests import oracle_lib as O


def main():
    out = sys.argv[1]
    w, h, frames = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (480, 270, 6)
    scene = scenes.small_city_scene_textured()
    scene.env_map = scenes.procedural_sky(128, 64)
    osc = O.OracleScene(scene)
    fr = O.OracleFrame(osc, w, h)
    p = abi.default_frame_params(scene, w, h)
    for f in range(frames):
        p.numAccumFrames = f
        for kind, pid in engine.restir_frame_passes(p, f, 1):
            fr.gbuffer(p) if kind == "gbuffer" else fr.restir(p, pid)
    rgba8 = O.present(fr.buffer(abi.BUF_BEAUTY_ACCUM), brightness_scale=1.0)
    imageio.write_png(out, rgba8)
    print(out, rgba8.shape)


if __name__ == "__main__":
    main()
