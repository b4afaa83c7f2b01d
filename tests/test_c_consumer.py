"""The drop-in boundary from plain C: include/jaicov_b200.h compiles as C99 (-pedantic -Werror), links against
libjaicov_b200.so alone and the INTEGRATION.md call sequence fails loudly without a device (no CPU path)."""
import os
import subprocess

import bundle_adjustment_b200 as ba

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c99_consumer(built, tmp_path):
    libdir = os.path.join(ROOT, 'bundle-adjustment_b200')
    exe = str(tmp_path / 'consumer')
    subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Wall', '-Wextra', '-Werror', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'c_consumer', 'consumer.c'), '-o', exe,
                           '-L', libdir, '-ljaicov_b200', '-Wl,-rpath,' + libdir])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'consumer ok' in r.stdout
    if ba._lib.load().jaicov_device_count() == 0:
        assert 'no sm_100 device' in r.stdout


def test_cpp_example_builds_and_runs(built, tmp_path):
    """examples/example_adjustment.cpp (the reference's example flow on the native C++ host mirror) compiles warning-free, indexes
    its network (bookkeeping runs on the host) and, without a device, stops at the library call with NOT_INITIALISED."""
    libdir = os.path.join(ROOT, 'bundle-adjustment_b200')
    exe = str(tmp_path / 'example_adjustment')
    subprocess.check_call(['g++', '-std=c++17', '-O1', '-Wall', '-Wextra', '-Werror', os.path.join(ROOT, 'examples', 'example_adjustment.cpp'),
                           '-o', exe, '-L', libdir, '-ljaicov_b200', '-Wl,-rpath,' + libdir])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'observations 1920, unknowns 314, datum defect 7, redundancy 1613' in r.stdout
    if ba._lib.load().jaicov_device_count() == 0:
        assert 'estimateModel() -> -5' in r.stdout
    else:
        assert 'sigma0 a posteriori / a priori' in r.stdout
