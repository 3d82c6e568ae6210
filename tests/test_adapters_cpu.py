"""the mp2p_icp adapters (adapters/mp2p_icp/) compile against the declarations of the upstream
API (tests/mock_upstream/, the real libraries are absent here) and behave at the reference's ICP seam
(/root/reference/src/LidarOdometry.cpp:57-88, 869-880) when driven against the device test double."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AD = os.path.join(ROOT, "adapters", "mp2p_icp")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    d = tmp_path_factory.mktemp("adapters")
    fake = str(d / "libfake_b200icp.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", fake, os.path.join(ROOT, "tests", "stub", "fake_b200icp.c")])
    exe = str(d / "adapter_driver")
    srcs = [os.path.join(AD, "src", f) for f in ("DeviceCloudCache.cpp", "ICP_B200.cpp", "Matcher_B200.cpp",
                                                  "FilterEdgesPlanes_B200.cpp", "register.cpp")]
    srcs += [os.path.join(ROOT, "tests", "mock_upstream", "mp2p_icp", "mock_register.cpp"),
             os.path.join(ROOT, "tests", "stub", "adapter_driver.cpp")]
    subprocess.check_call([CXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-pthread",
                           "-I", os.path.join(AD, "include"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "mock_upstream"), "-o", exe] + srcs +
                          [fake, "-Wl,-rpath," + str(d)])
    return exe


def test_adapters_compile_warning_free_and_drive_the_icp_seam(driver):
    out = subprocess.run([driver], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout
    assert "FAIL" not in out.stdout


def test_cmake_project_is_guarded_by_find_package():
    txt = open(os.path.join(AD, "CMakeLists.txt")).read()
    assert "find_package(mp2p_icp QUIET)" in txt and "return()" in txt
    for f in ("src/DeviceCloudCache.cpp", "src/ICP_B200.cpp", "src/Matcher_B200.cpp", "src/FilterEdgesPlanes_B200.cpp",
              "src/register.cpp"):
        assert f in txt and os.path.exists(os.path.join(AD, f))


def test_adapters_never_include_the_mock_or_the_oracle():
    for dirpath, _, files in os.walk(AD):
        for f in files:
            s = open(os.path.join(dirpath, f)).read()
            assert "mock_upstream" not in s.replace("tests/mock_upstream", "") or f == "README.md"
            assert "oracle" not in s
