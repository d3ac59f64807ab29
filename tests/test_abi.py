"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/zkb200.h declares, and refuses to
work without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ["zk_ctx_create", "zk_bases_upload", "zk_msm", "zk_msm_dev", "zk_msm_batch", "zk_ntt", "zk_ntt_batch", "zk_ntt_dev",
                 "zk_srs_commit_non_hiding", "zk_srs_commit_evaluations_non_hiding", "zk_srs_mask_custom", "zk_jacobian_to_affine"]:
        assert must in syms


def test_library_loads_and_exports_every_declared_symbol():
    import proof_systems_b200 as zk
    L = zk.lib()
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/zkb200.h but not exported by libzkb200.so"


def test_rust_shim_binds_only_what_the_header_declares():
    """crates/zkb200/src/ffi.rs (the reference-side binding, shipped as source: no Rust toolchain in this image) must name only
    functions the header declares and the library exports, with the header's argument counts; its constants must equal the header's"""
    ffi = open(os.path.join(ROOT, "crates", "zkb200", "src", "ffi.rs")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "zkb200.h")).read(), flags=re.S)
    rust = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (zk_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]*)?;", ffi, flags=re.S)}
    assert len(rust) >= 20
    declared = declared_symbols()
    import proof_systems_b200 as zk
    L = zk.lib()
    for name, args in rust.items():
        assert name in declared and hasattr(L, name), name
        c_args = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", hdr, flags=re.S).group(1)
        n_c = 0 if c_args.strip() in ("", "void") else c_args.count(",") + 1
        n_r = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        assert n_c == n_r, (name, n_c, n_r)
    for const, val in re.findall(r"pub const (ZK_[A-Z_]+): (?:c_int|u32) = (-?\d+);", ffi):
        m = re.search(r"#define\s+" + const + r"\s+\(?(-?\d+)\)?", hdr) or re.search(const + r"\s*=\s*(-?\d+)", hdr)
        assert m and int(m.group(1)) == int(val), const
    # every zk_* call in the crate's other files is declared in ffi.rs
    src_dir = os.path.join(ROOT, "crates", "zkb200", "src")
    for f in os.listdir(src_dir):
        if f.endswith(".rs") and f != "ffi.rs":
            code = re.sub(r"//.*", "", open(os.path.join(src_dir, f)).read())
            for called in set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", code)):
                assert called in rust, (f, called)
    # every source file of the crate is listed in lib.rs
    lib_rs = open(os.path.join(ROOT, "crates", "zkb200", "src", "lib.rs")).read()
    for f in os.listdir(os.path.join(ROOT, "crates", "zkb200", "src")):
        if f.endswith(".rs") and f != "lib.rs":
            assert f"pub mod {f[:-3]};" in lib_rs, f


def test_no_cpu_fallback():
    import torch
    import proof_systems_b200 as zk
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(zk.ZkError) as e:
        zk.Context(0)
    assert e.value.code == -3  # ZK_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "proof_systems_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pasta_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libzkb200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_header_is_plain_c_and_links(tmp_path):
    """include/zkb200.h is a C ABI: it must compile as C11 (-pedantic) and link against the library; without a GPU the first
    call fails loudly with ZK_ERR_NO_DEVICE."""
    import subprocess
    src = tmp_path / "abi_c.c"
    src.write_text('#include "zkb200.h"\n#include <stdio.h>\n'
                   'int main(void) { zk_ctx* ctx = NULL; int rc = zk_ctx_create(0, &ctx);\n'
                   '  printf("%d|%s\\n", rc, zk_last_error()); if (rc == ZK_OK) zk_ctx_destroy(ctx); return 0; }\n')
    exe = tmp_path / "abi_c"
    lib_dir = os.path.join(ROOT, "proof_systems_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                           "-L", lib_dir, "-lzkb200", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    rc = int(out.split("|")[0])
    import torch
    assert rc == (0 if torch.cuda.is_available() else -3), out


def test_cpp_host_layer_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    """include/zkb200.hpp (the C++ mirror of SRS<G> / Radix2EvaluationDomain / the open rounds) and its test driver compile
    warning-free; without a GPU the Context constructor throws zkb200::Error{ZK_ERR_NO_DEVICE}."""
    import subprocess
    import torch
    lib_dir = os.path.join(ROOT, "proof_systems_b200")
    exe = str(tmp_path / "host_layer")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_layer.cpp"), "-L", lib_dir, "-lzkb200", f"-Wl,-rpath,{lib_dir}", "-o", exe])
    if torch.cuda.is_available():
        return
    blob = tmp_path / "in.bin"
    np = __import__("numpy")
    np.zeros(1 + 8 * 4 + 8 + 6 * 4 + 4 * 4 + 2 * 4 + 8, dtype="<u8").tofile(str(blob))
    run = subprocess.run([exe, str(blob), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert run.returncode != 0 and not os.path.exists(str(tmp_path / "out.bin"))
