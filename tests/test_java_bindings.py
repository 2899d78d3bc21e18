"""CPU suite: the Java FFM binding (source only -- there is no JVM in the build image) is kept in sync with the C header
mechanically.  include/colq.h is parsed into FFM value layouts, ColqLibrary.java is parsed for its downcall handles, and
both are compared with each other and with the ctypes twin colq/_ffi.py that the GPU tests actually call through.
"""
import ctypes as C
import importlib.util
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
JAVA_DIR = ROOT / "java-columnar-query-engine_b200" / "java" / "data-system-b200" / "src"
spec = importlib.util.spec_from_file_location("gen_java_bindings", ROOT / "scripts" / "gen_java_bindings.py")
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def java_handles():
    text = gen.JAVA.read_text()
    out = {}
    for m in re.finditer(r'static final MethodHandle (colq_\w+) = h\("(colq_\w+)", ([A-Z_, ]+)\);', text):
        assert m.group(1) == m.group(2), "field name and symbol name differ"
        layouts = [x.strip() for x in m.group(3).split(",")]
        out[m.group(1)] = (layouts[0], layouts[1:])
    return out


def ctypes_layout(t):
    if t is None:
        return "VOID"
    if t in (C.c_int, C.c_int32):
        return "JAVA_INT"
    if t is C.c_int64:
        return "JAVA_LONG"
    return "ADDRESS"   # c_void_p, c_char_p, POINTER(...)


def test_every_header_symbol_is_bound_in_java_with_the_right_descriptor():
    header = {name: (ret, args) for name, ret, args in gen.parse_header()}
    java = java_handles()
    assert len(header) >= 60
    assert set(java) == set(header), (sorted(set(header) - set(java)), sorted(set(java) - set(header)))
    for name, sig in header.items():
        assert java[name] == sig, f"{name}: ColqLibrary.java binds {java[name]}, include/colq.h declares {sig}"


def test_generated_file_is_current():
    assert gen.JAVA.read_text() == gen.render(), "ColqLibrary.java is stale: run python scripts/gen_java_bindings.py"


def test_ctypes_twin_agrees_with_the_header():
    from colq import _ffi
    header = {name: (ret, args) for name, ret, args in gen.parse_header()}
    assert set(_ffi.SIGNATURES) == set(header)
    for name, (res, args) in _ffi.SIGNATURES.items():
        got = (ctypes_layout(res), [ctypes_layout(a) for a in args])
        assert got == header[name], f"{name}: colq/_ffi.py declares {got}, include/colq.h declares {header[name]}"


def test_shim_only_calls_bound_symbols_with_matching_arity():
    """Every `colq_xxx.invokeExact(...)` in the Java sources names a bound handle and passes as many arguments as its
    descriptor has (a wrong count would only show up at run time on a machine with a JVM)."""
    java = java_handles()
    n_calls = 0
    for path in JAVA_DIR.rglob("*.java"):
        text = path.read_text()
        for m in re.finditer(r"\b(colq_\w+)\.invokeExact\(", text):
            name = m.group(1)
            assert name in java, f"{path.name}: {name} is not bound in ColqLibrary.java"
            depth, i, n_args, any_arg = 1, m.end(), 0, False
            while depth:
                ch = text[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    n_args += 1
                elif not ch.isspace():
                    any_arg = True
                i += 1
            n_args = n_args + 1 if any_arg else 0
            assert n_args == len(java[name][1]), f"{path.name}: {name} called with {n_args} arguments, descriptor has {len(java[name][1])}"
            n_calls += 1
    assert n_calls > 50
