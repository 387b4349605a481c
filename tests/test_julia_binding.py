"""CPU: static cross-check of every `ccall` in the Julia wrapper and in INTEGRATION.md against the C
prototypes of include/maprast.h (symbol exists, return type, argument count, argument widths / pointee
types), of the ctypes binding's argtypes against the same prototypes, and a plain-C compile of the header.
Julia is not installed in the build image, so this is the only machine check the .jl file gets here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "maprast.h")
JL = os.path.join(ROOT, "maprasterization.jl_b200", "julia", "MapRasterizationB200.jl")
INTEG = os.path.join(ROOT, "INTEGRATION.md")


def c_class(t):
    """C type -> (kind, pointee)"""
    t = t.replace("const", " ").strip()
    t = re.sub(r"\s+", " ", t)
    if "*" in t:
        base = t.replace("*", " ").split()[0]
        return ("ptr", base)
    base = t.split()[0] if t.split() else "void"
    return ({"int": "i32", "int64_t": "i64", "double": "f64", "uint32_t": "u32", "size_t": "size", "void": "void",
             "float": "f32"}[base], None)


def prototypes():
    src = re.sub(r"/\*.*?\*/", " ", open(HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(mr_[a-z_0-9]+)\s*\(([^()]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                a = re.sub(r"\b[A-Za-z_]\w*$", "", a).strip() if not a.endswith("*") else a   # drop the parameter name
                arglist.append(c_class(a))
        protos[name] = (c_class(ret), arglist)
    return protos


JL_SCALAR = {"Cint": "i32", "Int32": "i32", "Int64": "i64", "Float64": "f64", "UInt32": "u32", "Cuint": "u32",
             "Csize_t": "size", "Cvoid": "void", "Cfloat": "f32"}
JL_POINTEE = {"UInt8": {"uint8_t"}, "Int64": {"int64_t"}, "Int32": {"int32_t"}, "Float64": {"double"}, "Cint": {"int"}, "Cfloat": {"float"},
              "MRCounters": {"mr_counters"}, "Cvoid": None}   # Ptr{Cvoid}: any pointer


def jl_class(t):
    t = t.strip()
    if t == "Cstring":
        return ("ptr", {"char"})
    m = re.fullmatch(r"(Ptr|Ref)\{(\w+)\}", t)
    if m:
        return ("ptr", JL_POINTEE[m.group(2)])
    return (JL_SCALAR[t], None)


def split_types(s):
    return [x.strip() for x in s.replace("\n", " ").split(",") if x.strip()]


def ccalls(path):
    txt = open(path).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libmaprast\),\s*([\w\{\}]+),\s*\(([^()]*)\)", txt, flags=re.S):
        out.append((m.group(1), m.group(2), split_types(m.group(3)), txt.count("\n", 0, m.start()) + 1))
    return out


@pytest.mark.parametrize("path", [JL, INTEG])
def test_every_ccall_matches_the_header(path):
    protos = prototypes()
    assert len(protos) >= 35
    calls = ccalls(path)
    assert calls, f"no ccall parsed in {path}"
    for name, ret, args, line in calls:
        where = f"{os.path.basename(path)}:{line} {name}"
        assert name in protos, f"{where}: not declared in maprast.h"
        cret, cargs = protos[name]
        jr = jl_class(ret)
        assert jr[0] == cret[0], f"{where}: return type {ret} vs C {cret}"
        assert len(args) == len(cargs), f"{where}: {len(args)} argument types, C prototype has {len(cargs)}"
        for k, (ja, ca) in enumerate(zip(args, cargs)):
            jk = jl_class(ja)
            assert jk[0] == ca[0], f"{where}: argument {k + 1} is {ja}, C has {ca}"
            if jk[0] == "ptr":
                assert jk[1] is None or ca[1] in jk[1], f"{where}: argument {k + 1} is {ja}, C pointee is {ca[1]}"


def test_julia_wrapper_covers_the_entry_points_it_documents():
    names = {c[0] for c in ccalls(JL)}
    for need in ("mr_create", "mr_destroy", "mr_last_error", "mr_set_palette", "mr_set_known", "mr_categorize_clean_host",
                 "mr_clean_host", "mr_clean_segments_host", "mr_fill_gaps_host", "mr_multi_create", "mr_multi_destroy",
                 "mr_multi_set_palette", "mr_multi_set_known", "mr_multi_categorize_clean_host",
                 "mr_blur_host", "mr_balance_host", "mr_categorize_clean_f64_host", "mr_stds_host", "mr_stripes_host",
                 "mr_set_texture_stats", "mr_categorize_clean_props_host", "mr_polygons_host", "mr_shrink_grow_host",
                 "mr_recolor_host"):
        assert need in names, f"the Julia wrapper never calls {need}"
    src = open(JL).read()
    # every entry point that runs with a known plane clears it again (stale points)
    assert src.count("_clear_known(ctx)") >= 4
    assert "applicable(rebuild" not in src


def test_ctypes_argtypes_match_the_header(mr):
    protos = prototypes()
    L = mr._lib.load()
    width = {"i32": 4, "i64": 8, "f64": 8, "u32": 4, "f32": 4, "size": C.sizeof(C.c_size_t)}
    for name, (cret, cargs) in protos.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            assert not cargs or name in ("mr_version",), f"{name}: no argtypes set in _lib.py"
            continue
        assert len(fn.argtypes) == len(cargs), f"{name}: {len(fn.argtypes)} argtypes, C prototype has {len(cargs)}"
        for k, (pt, ca) in enumerate(zip(fn.argtypes, cargs)):
            is_ptr = pt in (C.c_void_p, C.c_char_p) or hasattr(pt, "contents") or hasattr(pt, "_type_") and isinstance(pt._type_, type)
            if ca[0] == "ptr":
                assert is_ptr, f"{name}: argument {k + 1} should be a pointer, is {pt}"
            else:
                assert not is_ptr and C.sizeof(pt) == width[ca[0]], f"{name}: argument {k + 1} is {pt}, C has {ca[0]}"
                if ca[0] == "f64":
                    assert pt is C.c_double


def test_header_compiles_as_plain_c():
    protos = prototypes()
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        with open(src, "w") as f:
            f.write('#include "maprast.h"\n')
            f.write("void* table[] = {\n" + "".join(f"    (void*){n},\n" for n in sorted(protos)) + "};\n")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-Wno-pedantic", "-c", "-I",
                               os.path.join(ROOT, "include"), src, "-o", os.path.join(d, "t.o")])
