"""oracle/nrc.cpp's hash-grid and one-blob encodings against the reference's own tiny-cuda-nn kernels (kernel_grid<__half,3,2>,
kernel_one_blob_soa<__half>) compiled from /root/reference/ext/tiny-cuda-nn into oracle/_ref/libtcnn_ref.so with -fmad=false, i.e.
under the oracle's arithmetic model (the reference's own build contracts a*b+c into FMAs, which no CPU restatement can follow
bit for bit; what is compared here is the algorithm): every half bit-identical.  First passed on the round-1 driver box (XPASS);
a plain test since round 2.  The check runs in a child process so that a fault inside third-party kernels cannot poison this
session's CUDA context."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_encoding_equals_tcnn_kernels():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libtcnn_ref.so")):
        pytest.skip("oracle/_ref/libtcnn_ref.so not built (needs /root/reference at build time)")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tcnn_ref_check.py")], capture_output=True, text=True,
                          timeout=240)
    line = proc.stdout.strip().splitlines()[-1] if proc.stdout.strip() else ""
    assert proc.returncode == 0, f"rc={proc.returncode} {line} {proc.stderr[-400:]}"
    result = json.loads(line)
    assert result["grid"]["mismatches"] == 0 and result["oneblob"]["mismatches"] == 0, line
