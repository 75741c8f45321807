"""The bounding-sphere pre-test of the light-triangle fetch (step 0 of sampleLightUnlessDark, gfxexp_b200/csrc/lighting.cuh; on since round 2)
never rejects a triangle that has a sample point above - or within the 1e-3 cosine margin of - the shading horizon:
3 M random configurations on the CPU, in double precision."""
import os
import subprocess


def test_sphere_pretest_is_conservative(tmp_path):
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "sphere_cull_check.cpp")
    exe = str(tmp_path / "sphere_cull_check")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-o", exe, src], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip().endswith(" 0 violations"), out
