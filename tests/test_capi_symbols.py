"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(caae_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_hot_path_entry_points():
    names = _declared_symbols()
    for required in ("caae_fps", "caae_gather", "caae_gather_grad", "caae_nn_distance", "caae_nn_distance_grad",
                     "caae_prob_sample"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from cloudaae_b200 import _capi
    lib = _capi.lib()  # raises loudly when the .so is missing
    handle = ctypes.CDLL(_capi.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/ but not exported"
    assert lib.caae_abi_version() == _capi.ABI_VERSION
    assert set(_capi.EXPORTED_SYMBOLS) <= set(_declared_symbols())


def test_status_strings_and_argument_errors_need_no_gpu():
    from cloudaae_b200 import _capi
    lib = _capi.lib()
    assert lib.caae_status_string(0) == b"ok"
    assert b"scratch" in lib.caae_status_string(-3)
    # argument validation happens before any CUDA call
    assert lib.caae_nn_distance(-1, 1, None, 1, None, None, None, None, None, None) == -1
    assert lib.caae_nn_distance(1, 4, None, 4, None, None, None, None, None, None) == -2
    assert lib.caae_fps(1, 0, 4, None, None, None, None) == -1
    assert lib.caae_fps(0, 0, 4, None, None, None, None) == 0
    assert lib.caae_fps_scratch_bytes(4, 2048) == 0
    assert lib.caae_fps_scratch_bytes(4, 10000) == 4 * 10000 * 4
    with pytest.raises(_capi.CloudAAENativeError, match="null pointer"):
        _capi.check(-2, "x")


def test_product_package_never_imports_the_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "cloudaae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "caae_oracle" not in text and "libcloudaae_ref" not in text, f


def _header_prototypes():
    text = open(os.path.join(ROOT, "include", "cloudaae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|size_t|const char\*)\s+(caae_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        codes = ""
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a or a.startswith("caae_stream_t"):
                    codes += "p"
                elif a.startswith("int "):
                    codes += "i"
                elif a.startswith("unsigned long long "):
                    codes += "Q"
                elif a.startswith("unsigned int "):
                    codes += "I"
                elif a.startswith("long "):
                    codes += "l"
                elif a.startswith("float "):
                    codes += "f"
                elif a.startswith("double "):
                    codes += "d"
                else:
                    raise AssertionError(f"unparsed argument {a!r} in {name}")
        protos[name] = codes
    return protos


def test_ctypes_signatures_match_the_header():
    import ctypes as C
    from cloudaae_b200 import _capi
    back = {C.c_int: "i", C.c_long: "l", C.c_float: "f", C.c_double: "d", C.c_void_p: "p", C.c_ulonglong: "Q"}
    protos = _header_prototypes()
    for name, argtypes in _capi._SIGNATURES.items():
        got = "".join(back[a] for a in argtypes)
        assert name in protos, name
        assert got == protos[name], f"{name}: ctypes {got} vs header {protos[name]}"
