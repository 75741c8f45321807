"""host/png_writer.h (the C++ host's stand-in for the reference's stb writer) produces files that a PNG decoder accepts and
that hold exactly the packed RGBA8 image: CPU test program + zlib."""
import os
import subprocess

import numpy as np

from gfxexp_b200 import imageio


def test_cpp_png_writer_round_trip(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "png_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tests", "cpp", "png_check.cpp")], check=True)
    for w, h in ((7, 5), (300, 250)):      # the second image needs several 64 KB stored blocks
        path = str(tmp_path / f"t{w}.png")
        subprocess.run([exe, path, str(w), str(h)], check=True)
        s = np.uint32(12345)
        want = np.empty(w * h, dtype=np.uint32)
        state = 12345
        for i in range(w * h):
            state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
            want[i] = state
        got = imageio.read_png(path)
        assert got.shape == (h, w) and np.array_equal(got.reshape(-1), want)
