"""CPU: the C-ABI library builds, loads and exports every symbol include/mtvaf_b200.h declares
(no compute calls without a GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mtvaf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtvaf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mtvaf_b200 import build
    build.build()
    from mtvaf_b200 import lib
    names = header_functions()
    assert len(names) >= 20
    raw = lib.raw()
    for n in names:
        assert hasattr(raw, n), "symbol %s declared in the header but not exported" % n
    # and the ctypes table binds every one of them
    for n in names:
        if n not in ("mtvaf_last_error", "mtvaf_launch_count", "mtvaf_attention_fwd_workspace_bytes"):   # non-int returns
            assert n in lib.SIGNATURES, n
    hdr = open(os.path.join(ROOT, "include", "mtvaf_b200.h")).read()
    assert lib.abi_version() == int(re.search(r"#define MTVAF_ABI_VERSION (\d+)", hdr).group(1)) == 3
    # the ctypes mirror of MtvafEpilogue must have the header's fields, in order
    body = re.search(r"typedef struct MtvafEpilogue \{(.*?)\} MtvafEpilogue;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(\w+)\s*;", body)
    assert fields == [f[0] for f in lib.Epilogue._fields_], (fields, lib.Epilogue._fields_)


def test_no_cpu_fallback_in_product_package():
    """The product package must never import the oracle."""
    pkg = os.path.join(ROOT, "mtvaf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                s = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in s and "from oracle" not in s, os.path.join(dirpath, f)


def test_bad_arguments_return_error_not_crash():
    import ctypes as C
    from mtvaf_b200 import lib
    ep = lib.Epilogue()
    rc = lib.raw().mtvaf_gemm_bf16(None, 0, 0, None, 0, 0, 0, 0, 0, C.byref(ep), 1, None)
    assert rc == -1 and "null" in lib.last_error()
    rc = lib.raw().mtvaf_probe_labels(None, None, 0, 0, None)
    assert rc == -1


def test_header_is_valid_c99_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/mtvaf_b200.h must compile as plain C (and C++), and a C host must link against
    the shared library and read the ABI version (no compute: no GPU needed)."""
    import subprocess
    from mtvaf_b200 import build
    lib_path = build.build()
    src = tmp_path / "host.c"
    src.write_text('#include "mtvaf_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { MtvafEpilogue e; e.mode = MTVAF_EPI_STORE; e.colsum = 0; (void)e;\n'
                   '  printf("%d %d\\n", mtvaf_abi_version(), MTVAF_ABI_VERSION); return 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++17")):
        r = subprocess.run([cc, std, "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-I", inc] +
                           (["-x", "c++"] if cc == "g++" else []) + [str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    exe = tmp_path / "host"
    r = subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), lib_path,
                        "-Wl,-rpath," + os.path.dirname(lib_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = r.stdout.split()
    assert a == b == "3"
