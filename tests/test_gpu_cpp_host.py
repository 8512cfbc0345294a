"""The reference's own hot-path unit tests, restated in C++ on the host mirror of the Rust API
(baby-plonk-rust_b200/host/baby_plonk.hpp -> C ABI -> CUDA): tests/cpp/reference_style_tests.cpp."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _exe():
    import __graft_entry__ as ge
    return ge.build_cpp_host_tests()


@pytest.mark.gpu
def test_reference_style_cpp_tests_pass_on_gpu():
    r = subprocess.run([_exe()], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASSED" in r.stdout
    for name in ("test_generate_srs", "test_monomial_commit", "test_ntt_round_trip", "test_polynomial_mul",
                 "test_root_of_unity", "test_bucket_msm"):
        assert name + " ... ok" in r.stdout


def test_cpp_host_layer_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_exe()], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stdout
