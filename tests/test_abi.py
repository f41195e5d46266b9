"""CPU tests of the boundary: libflinkms_b200.so loads without a GPU, exports exactly the symbols include/flinkms.h
declares, the ctypes table covers them all, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flinkms.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"FMS_API\s+[\w\s\*]+?\b(fms_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("fms_create", "fms_destroy", "fms_last_error", "fms_als_fit", "fms_als_get_factors", "fms_als_predict",
                 "fms_svm_fit", "fms_svm_predict", "fms_als_half_step", "fms_svm_sweep"):
        assert must in names


def test_library_exports_every_declared_symbol(fms):
    lib = fms._native.load()
    names = _declared()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in flinkms.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", fms._native.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r"\bT (fms_\w+)", out)))
    assert exported == names, "exported fms_* symbols and flinkms.h drifted apart"
    assert sorted(fms._native.SIGNATURES) == names, "ctypes signature table and flinkms.h drifted apart"


def test_library_has_no_cuda_needed_entries(fms):
    out = subprocess.check_output(["readelf", "-d", fms._native.LIB_PATH], text=True)
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not [n for n in needed if "cuda" in n or "cudart" in n or "nccl" in n or "oracle" in n], needed


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "flink-ms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".c", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_no_gpu_means_loud_failure(fms):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fms.FlinkMSError) as e:
        fms.Context(0)
    assert e.value.status == fms._native.ERR_CUDA and "no CPU fallback" in str(e.value)
    lib = fms._native.load()
    assert lib.fms_version() >= 100
    # null ctx: argument error, not a crash
    assert lib.fms_als_half_step(None, 0, ctypes.c_double(0.1)) == fms._native.ERR_INVALID_ARG
    assert lib.fms_create(None, None, 1) == fms._native.ERR_INVALID_ARG
    h = ctypes.c_void_p()
    assert lib.fms_create(ctypes.byref(h), None, 2) == fms._native.ERR_INVALID_ARG      # a group needs explicit device ids


def test_jni_shim_type_checks_against_the_header_and_matches_native_java():
    """No JDK in the build image: the JNI shim (flink-ms_b200/jvm/jni/flinkms_jni.c) is source-only.  It still must (a) type-check
    against include/flinkms.h -- with a stand-in <jni.h> that declares just the JNI calls the shim uses -- and (b) implement
    exactly the native methods Native.java declares."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(root, "flink-ms_b200", "jvm", "jni", "flinkms_jni.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "tests", "jni_stub"), "-I",
                        os.path.join(root, "include"), shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    c_names = sorted(set(re.findall(r"Java_de_tub_it4bi_gpu_Native_(\w+)", open(shim).read())))
    java = open(os.path.join(root, "flink-ms_b200", "jvm", "src", "main", "java", "de", "tub", "it4bi", "gpu", "Native.java")).read()
    j_names = sorted(set(re.findall(r"private static native [\w\[\]]+ (\w+)\(", java)))
    assert c_names == j_names
