"""The C-ABI library loads and exports exactly the entry points declared in include/bt_api.h (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bt_api.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bt_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from brax_tracking_b200 import build, native
    lib = build.build()
    assert os.path.exists(lib)
    l = ctypes.CDLL(lib)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(l, s), f"{s} is declared in include/bt_api.h but not exported"
    assert sorted(native.SYMBOLS) == syms, "native.SYMBOLS and include/bt_api.h disagree"


def test_missing_library_fails_loudly(monkeypatch):
    from brax_tracking_b200 import native
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libbt_b200.so")
    try:
        native.lib()
    except RuntimeError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("a missing CUDA library must raise")


def test_error_path_without_gpu():
    """bt_model_create with bad arguments returns an error code and a message instead of crashing."""
    from brax_tracking_b200 import native
    l = native.lib()
    rc = l.bt_model_create(0, None, None, None, None, 0, None)
    assert rc < 0 and b"null" in l.bt_last_error()
    assert l.bt_launch_count() >= 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "brax_tracking_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".inc", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|env_oracle|emu)\b", text, flags=re.M), f
                assert "liboracle" not in text and "libbt_emu" not in text, f


def test_xla_ffi_shim_type_checks(tmp_path):
    """csrc/xla_ffi_shim.cc (BtStepFfi, BtResetFfi) against a stand-in of jaxlib's typed-FFI binder (tests/ffi_mock): every
    handler is invocable with exactly the types its binding declares -- 19 operands (9 of them aliased in/out state) + 12 results
    for the step, 1 operand + 13 results for the reset -- and both symbols are emitted."""
    import subprocess
    so = str(tmp_path / "shim.so")
    subprocess.run(["g++", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "tests", "ffi_mock"), "-I" + os.path.join(ROOT, "include"),
                    "-I/usr/local/cuda/include", os.path.join(ROOT, "brax_tracking_b200", "csrc", "xla_ffi_shim.cc"), "-o", so,
                    "-Wl,--unresolved-symbols=ignore-all"], check=True)
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    assert "BtStepFfi_arity" in out and "BtResetFfi_arity" in out
    src = open(os.path.join(ROOT, "brax_tracking_b200", "csrc", "xla_ffi_shim.cc")).read()
    step = src[src.index("XLA_FFI_DEFINE_HANDLER_SYMBOL(BtStepFfi"):src.index("// wrap(env).reset")]
    assert step.count(".Arg<") == 19 and step.count(".Ret<") == 12
    reset = src[src.index("XLA_FFI_DEFINE_HANDLER_SYMBOL(BtResetFfi"):]
    assert reset.count(".Arg<") == 1 and reset.count(".Ret<") == 13
