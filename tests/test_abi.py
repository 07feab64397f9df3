"""The C-ABI library builds in-tree, loads, and exports every symbol include/basd_b200.h
declares (no compute here: there is no GPU in the build container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "basd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long)\s+(basd_\w+)\s*\(", text)))


def test_library_exports_header_symbols():
    import __graft_entry__ as entry
    entry.build()
    from basd_b200 import _native as nat
    lib = ctypes.CDLL(nat.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in basd_b200.h but not exported"


def test_binding_table_matches_header():
    from basd_b200 import _native as nat
    declared = set(header_symbols())
    bound = set(nat.exported_symbols())
    assert bound <= declared, f"bound but undeclared: {bound - declared}"
    assert declared - bound <= set(nat._OPTIONAL), f"declared but unbound: {declared - bound}"


def test_missing_library_fails_loudly(monkeypatch):
    from basd_b200 import _native as nat
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libbasd_b200.so")
    try:
        nat.load()
    except nat.NativeLibraryMissing as e:
        assert "no" in str(e) and "fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")


def test_cpu_tensors_are_rejected():
    import torch
    import types
    from basd_b200.losses import BASDLoss
    mod = BASDLoss(torch.nn.CrossEntropyLoss(), 192, 384, 12, 64,
                   config=types.SimpleNamespace(num_extraction_points=4), teacher_has_cls_token=True)
    st = {l: torch.randn(16, 64, 192) for l in mod.token_layers}
    te = {l: torch.randn(16, 64, 384) for l in range(2)}
    at = {l: torch.softmax(torch.randn(16, 2, 65, 65), -1) for l in range(2)}
    try:
        mod(torch.randn(16, 10), torch.zeros(16, dtype=torch.long), st, te, at)
    except RuntimeError as e:
        assert "CUDA" in str(e) or "CPU" in str(e)
    else:
        raise AssertionError("CPU tensors must not silently run")


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/basd_b200.h is consumed as C (not C++) and a C program links against the library:
    the host-only sizing entry points answer without a GPU (what a cgo / JNI / plain-C binding
    would call first)."""
    import shutil
    import subprocess
    import __graft_entry__ as entry
    entry.build()
    from basd_b200 import _native as nat
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "basd_b200.h")], check=True)
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include "basd_b200.h"\n'
        "int main(void) {\n"
        '  printf("%d %ld %d\\n", basd_weight_grad_slices(),\n'
        "         basd_token_gram_tc_workspace_bytes(50176L, 768),\n"
        "         basd_gemm_tc3_supported(196, 196, 196, 196, 196, 196, 38416L, 38416L, 38416L));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "probe"
    libdir = os.path.dirname(nat.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lbasd_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    slices, ws_bytes, ok = int(out[0]), int(out[1]), int(out[2])
    assert slices > 0 and slices % 148 == 0              # grid sized in multiples of the SM count
    assert ws_bytes >= 768 * 768 * 4                     # at least one D x D fp32 partial
    assert ok in (0, 1)


def header_prototypes():
    """name -> list of ctypes kinds ('p', 'i', 'l', 'f', 'd') parsed from the header's prototypes."""
    text = open(os.path.join(ROOT, "include", "basd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|long)\s+(basd_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        name, params = m.group(1), " ".join(m.group(2).split())
        kinds = []
        if params not in ("", "void"):
            for prm in params.split(","):
                prm = prm.strip()
                if "*" in prm:
                    kinds.append("p")
                elif re.match(r"(const\s+)?double\b", prm):
                    kinds.append("d")
                elif re.match(r"(const\s+)?float\b", prm):
                    kinds.append("f")
                elif re.match(r"(const\s+)?long\b", prm):
                    kinds.append("l")
                elif re.match(r"(const\s+)?int\b", prm):
                    kinds.append("i")
                else:
                    raise AssertionError(f"{name}: parameter type not understood: {prm!r}")
        protos[name] = kinds
    return protos


def test_binding_argument_types_match_the_header():
    """Every ctypes signature in _native.py has the arity and the scalar kinds of the prototype in
    include/basd_b200.h (a drifted binding would corrupt a call silently: ctypes checks nothing)."""
    from basd_b200 import _native as nat
    kind = {ctypes.c_void_p: "p", ctypes.c_int: "i", ctypes.c_long: "l", ctypes.c_float: "f",
            ctypes.c_double: "d"}
    protos = header_prototypes()
    table = dict(nat._SIG)
    table.update({k: v[0] for k, v in nat._OPTIONAL.items()})
    assert len(table) >= 40
    for name, argtypes in table.items():
        assert name in protos, f"{name} bound but no prototype parsed"
        got = [kind[t] for t in argtypes]
        assert got == protos[name], f"{name}: binding {got} != header {protos[name]}"
