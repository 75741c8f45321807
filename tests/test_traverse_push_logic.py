"""The branch-free push of traverse.cuh (-DGFX_TRAVERSE_PREDICATED_PUSH, switched off until it has been measured on a GPU)
leaves the same stack, the same next node and the same overflow state as the loop it replaces: 4 M random nodes on the CPU."""
import os
import subprocess


def test_predicated_push_equals_loop(tmp_path):
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "push_check.cpp")
    exe = str(tmp_path / "push_check")
    subprocess.run(["g++", "-O2", "-o", exe, src], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip().endswith("0 mismatches"), out
