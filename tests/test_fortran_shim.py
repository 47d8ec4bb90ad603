"""Static cross-check of the ISO_C_BINDING shim (musubi_b200/fortran/mus_b200_module.f90) against
include/musb200.h -- no Fortran compiler exists in this image, so the interface blocks are parsed
here: every bound name is declared in the header with the same number of arguments, by-value /
by-reference passing agrees with scalar / pointer parameters, and the C types agree with the kinds."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "musb200.h")
SHIM = os.path.join(ROOT, "musubi_b200", "fortran", "mus_b200_module.f90")

C_KIND = {"int": "c_int", "double": "c_double", "int32_t": "c_int32_t", "int64_t": "c_int64_t",
          "long long": "c_long_long", "size_t": "c_size_t", "char": "c_char", "void": "c_ptr",
          "unsigned long long": "c_long_long"}


def c_prototypes():
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(musb200_\w+)\s*\(([^)]*)\)\s*;", text):
        params = []
        body = " ".join(m.group(2).split())
        if body and body != "void":
            for p in body.split(","):
                p = p.replace("const", " ").strip()
                pointer = p.count("*")
                base = re.sub(r"\*", " ", p)
                base = " ".join(base.split()[:-1]) if len(base.split()) > 1 else base.strip()
                params.append((base.strip(), pointer))
        protos[m.group(1)] = params
    return protos


def fortran_interfaces():
    src = open(SHIM).read()
    src = re.sub(r"&\s*\n\s*&", " ", src)                       # join continuation lines
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)\s*result\(rc\)(.*?)end function",
                         src, flags=re.S | re.I):
        fname, args, cname, body = m.group(1), m.group(2), m.group(3), m.group(4)
        args = [a.strip() for a in args.split(",") if a.strip()]
        decl = {}
        for line in body.splitlines():
            line = line.split("!")[0].strip()
            if "::" not in line or line.lower().startswith("import"):
                continue
            spec, names = line.split("::")
            kind = re.search(r"(c_\w+)", spec)
            value = bool(re.search(r"\bvalue\b", spec, flags=re.I))
            for n in re.split(r",(?![^()]*\))", names):
                n = re.sub(r"\(.*\)", "", n).strip()
                decl[n.lower()] = (kind.group(1).lower() if kind else None, value)
        out[cname] = (fname, args, decl)
    return out


def test_header_parses_to_the_52_entry_points():
    protos = c_prototypes()
    assert len(protos) == 52 and "musb200_step" in protos
    assert protos["musb200_step"] == [("int", 0), ("int", 0), ("int", 0)]
    assert protos["musb200_finalize"] == []


def test_every_shim_binding_matches_the_header():
    protos, ifc = c_prototypes(), fortran_interfaces()
    assert len(ifc) >= 38
    for cname, (fname, args, decl) in ifc.items():
        assert fname == cname, "bind name %s differs from the interface name %s" % (cname, fname)
        assert cname in protos, "%s is not declared in include/musb200.h" % cname
        params = protos[cname]
        assert len(args) == len(params), "%s: %d dummy arguments, %d C parameters" % (cname, len(args), len(params))
        for a, (ctype, pointer) in zip(args, params):
            assert a.lower() in decl, "%s: dummy argument %s has no declaration" % (cname, a)
            kind, value = decl[a.lower()]
            want = C_KIND[ctype]
            if pointer == 0:
                assert value and kind == want, "%s(%s): C passes %s by value, shim declares %s%s" % (
                    cname, a, ctype, kind, ", value" if value else " by reference")
            elif kind == "c_ptr":
                assert value, "%s(%s): type(c_ptr) must be passed by value for a C pointer" % (cname, a)
            else:
                assert not value, "%s(%s): C takes a pointer, shim passes by value" % (cname, a)
                # void* takes any array by reference; typed pointers need the matching kind
                assert pointer == 1 and (ctype == "void" or kind == want), \
                    "%s(%s): C %s*, shim kind %s" % (cname, a, ctype, kind)
        assert decl.get("rc", (None, None))[0] == "c_int"


def test_every_bound_function_is_used_by_a_wrapper():
    src = open(SHIM).read()
    body = src[src.lower().index("contains"):]
    for cname in fortran_interfaces():
        assert re.search(r"\b%s\s*\(" % cname, body), "%s is bound but never called" % cname


def test_every_derived_type_component_the_shim_touches_exists_in_the_reference():
    """No compiler can check the shim here, so at least every `%component` it dereferences must
    be a component (or type-bound procedure) declared somewhere in the reference's sources."""
    import pytest
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree absent")
    src = re.sub(r"!.*", "", open(SHIM).read())
    used = set(m.lower() for m in re.findall(r"%\s*([A-Za-z_]\w*)", src))
    assert len(used) > 40
    declared = set()
    for base in (os.path.join(ref, "mus", "source"), os.path.join(ref, "tem", "source")):
        for d, _, files in os.walk(base):
            for f in files:
                if not f.endswith((".f90", ".fpp", ".inc")):
                    continue
                text = open(os.path.join(d, f), errors="replace").read()
                text = re.sub(r"&\s*\n\s*&?", " ", text)
                for line in text.splitlines():
                    line = line.split("!")[0]
                    if "::" in line:
                        for n in re.split(r",(?![^()]*\))", line.split("::", 1)[1]):
                            n = re.sub(r"\(.*|=.*", "", n).strip().lower()
                            if n:
                                declared.add(n)
                    m = re.match(r"\s*procedure\b.*?(\w+)\s*(=>.*)?$", line, flags=re.I)
                    if m:
                        declared.add(m.group(1).lower())
    # tem_communication_type%buf_real is generated by a CoCo text macro (buf_?tname?,
    # tem_comm_module.fpp:148-154) and used as me%buf_real throughout that file
    declared.add("buf_real")
    missing = sorted(used - declared)
    assert not missing, "components not found in the reference: %s" % missing
