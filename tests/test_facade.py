"""The C++ facade (cpppathtracer_b200/facade/*.h: the reference's class names over the C ABI).
CPU: it compiles against the header-only facade, the host camera controls reproduce the reference's
MotionalCamera bit for bit (golden poses), and without a device PathTracer reports the error instead of
falling back.  GPU: the VideoRenderer-style driver runs end to end."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden

LIBDIR = os.path.join(ROOT, "cpppathtracer_b200")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _compile(src, out):
    subprocess.check_call([CXX, "-std=c++17", "-O1", "-Wall", "-ffp-contract=off", "-o", str(out), os.path.join(ROOT, "tests", "facade", src),
                           f"-L{LIBDIR}", "-lpt_b200", f"-Wl,-rpath,{LIBDIR}", "-lpthread"])


def test_facade_camera_controls_match_reference(tmp_path):
    exe = tmp_path / "camera_ops"
    _compile("camera_ops.cpp", exe)
    g = golden("sky_camera.npz")
    lines = "".join(f"{int(op)} {float(arg)!r}\n" for op, arg in g["ops"])
    # the golden args were float32 values; print them exactly
    lines = "".join(f"{int(op)} {np.float32(arg):.9g}\n" for op, arg in g["ops"])
    out = subprocess.run([str(exe)], input=lines, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    poses = g["poses"][1:]
    assert len(out) == len(poses) + 1
    for line, want in zip(out, poses):
        got = np.array([struct.unpack("<f", struct.pack("<I", int(h, 16)))[0] for h in line.split()], np.float32)
        assert np.array_equal(got, want), (got, want)
    assert out[-1].endswith("idx 1")


def test_facade_driver_without_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    exe = tmp_path / "test_facade"
    _compile("test_facade.cpp", exe)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3 and r.stdout.startswith("NO_DEVICE")


@pytest.mark.gpu
def test_facade_driver_end_to_end(tmp_path):
    exe = tmp_path / "test_facade"
    _compile("test_facade.cpp", exe)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.splitlines()
    assert out[0] == "ids 0 1 2 3"
    f = out[1].split()
    assert f[1] == "4" and f[2] == "320x176" and f[4] == "4" and "err ''" in out[1]
    assert 255 * 320 * 176 < int(f[6]) < 4 * 255 * 320 * 176           # alpha is 255 everywhere, colours are not black
    assert out[2].startswith("after_move idx 1 changed 1")
    assert out[3] == "after_update changed 1"
    assert out[4] == "threaded frames 7"
    # pipelined DispatchRay: no callback after the first dispatch, one per dispatch afterwards, the last at Flush
    assert out[5] == "pipelined 0 1 2 3 err ''"


def _compile_c(src, out):
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-o", str(out), os.path.join(ROOT, "tests", "facade", src),
                           f"-L{LIBDIR}", "-lpt_b200", f"-Wl,-rpath,{LIBDIR}"])


def test_abi_header_is_plain_c(tmp_path):
    """include/pt_api.h compiles as pedantic C99 and links from C; without a device the example reports it and stops."""
    import torch
    exe = tmp_path / "abi_c_example"
    _compile_c("abi_c_example.c", exe)
    if torch.cuda.is_available():
        pytest.skip("device run is test_abi_c_example_runs")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3 and r.stdout.startswith("NO_DEVICE")


@pytest.mark.gpu
def test_abi_c_example_runs(tmp_path):
    exe = tmp_path / "abi_c_example"
    _compile_c("abi_c_example.c", exe)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK samples 4 mean 0.") and " moved 1 " in r.stdout, r.stdout + r.stderr
