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


def test_header_parses_to_the_entry_points_the_binding_knows():
    from musubi_b200._lib import SIGNATURES
    protos = c_prototypes()
    assert set(protos) == set(SIGNATURES) and len(protos) >= 55
    for name, params in protos.items():
        assert len(params) == len(SIGNATURES[name]), name
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


# ---------------------------------------------------------------------------------------------
# a type-aware pass: every component chain root%a(...)%b%c of the shim is resolved through the
# derived-type definitions of the reference (a compiler's job; none is available here)
def _reference_types(ref):
    """{type name: {component: type name or None (intrinsic / procedure)}} incl. inherited ones"""
    types, parents = {}, {}
    for base in (os.path.join(ref, "mus", "source"), os.path.join(ref, "tem", "source")):
        for d, _, files in os.walk(base):
            for f in files:
                if not f.endswith((".f90", ".fpp", ".inc")):
                    continue
                text = open(os.path.join(d, f), errors="replace").read()
                text = re.sub(r"&\s*\n\s*&?", " ", text)
                cur = None
                for raw in text.splitlines():
                    line = raw.split("!")[0].strip()
                    low = line.lower()
                    m = re.match(r"type\s*(?:,\s*([^:]*?))?\s*(?:::)?\s*(\w+)\s*$", line, flags=re.I)
                    if cur is None and m and not low.startswith("type("):
                        cur = m.group(2).lower()
                        types.setdefault(cur, {})
                        ext = re.search(r"extends\s*\(\s*(\w+)\s*\)", m.group(1) or "", flags=re.I)
                        if ext:
                            parents[cur] = ext.group(1).lower()
                        continue
                    if cur is not None:
                        if re.match(r"end\s*type", low):
                            cur = None
                            continue
                        if low.startswith("contains"):
                            continue
                        if "::" in line:
                            spec, names = line.split("::", 1)
                            t = re.match(r"\s*(?:type|class)\s*\(\s*(\w+)\s*\)", spec, flags=re.I)
                            tname = t.group(1).lower() if t else None
                            for n in re.split(r",(?![^()]*\))", names):
                                n = re.sub(r"\(.*|=.*", "", n).strip().lower()
                                if n:
                                    types[cur][n] = tname
                        else:
                            m2 = re.match(r"procedure\b.*?(\w+)\s*(=>.*)?$", line, flags=re.I)
                            if m2:
                                types[cur][m2.group(1).lower()] = None
    for t in list(types):
        p = parents.get(t)
        while p:
            for k, v in types.get(p, {}).items():
                types[t].setdefault(k, v)
            p = parents.get(p)
    # containers generated by CoCo text macros (tem_grow_array.fpp / tem_dyn_array.fpp, ?tname?):
    for name in ("grw_intarray_type", "grw_longarray_type", "grw_realarray_type", "grw_int2darray_type",
                 "grw_logical2darray_type", "dyn_intarray_type", "dyn_longarray_type"):
        types.setdefault(name, {}).update({"nvals": None, "val": None, "containersize": None, "sorted": None})
    # tem_communication_type%buf_real / tem_realbuffer_type come from the same kind of macro
    types.setdefault("tem_communication_type", {})["buf_real"] = "tem_realbuffer_type"
    types.setdefault("tem_realbuffer_type", {}).update({"pos": None, "nvals": None, "val": None})
    return types


def _chains(expr_text):
    """component chains 'a%b(..)%c' of a source text -> lists of names"""
    out, i, n = [], 0, len(expr_text)
    while i < n:
        m = re.compile(r"[A-Za-z_]\w*").match(expr_text, i)
        if not m or (i > 0 and (expr_text[i - 1].isalnum() or expr_text[i - 1] in "_%")):
            i += 1
            continue
        names, j = [m.group(0).lower()], m.end()
        while True:
            k = j
            while k < n and expr_text[k] == " ":
                k += 1
            if k < n and expr_text[k] == "(":                    # skip a balanced argument list
                depth = 0
                while k < n:
                    depth += expr_text[k] == "("
                    depth -= expr_text[k] == ")"
                    k += 1
                    if depth == 0:
                        break
                while k < n and expr_text[k] == " ":
                    k += 1
            if k < n and expr_text[k] == "%":
                m2 = re.compile(r"\s*([A-Za-z_]\w*)").match(expr_text, k + 1)
                if not m2:
                    break
                names.append(m2.group(1).lower())
                j = m2.end()
                continue
            break
        if len(names) > 1:
            out.append(names)
        i = m.end()
    return out


def test_component_chains_resolve_through_the_references_type_definitions():
    import pytest
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree absent")
    types = _reference_types(ref)
    assert "mus_scheme_type" in types and "layout" in types["mus_scheme_type"] and len(types) > 200
    src = re.sub(r"&\s*\n\s*&?", " ", re.sub(r"!.*", "", open(SHIM).read()))
    body = src[src.lower().index("\ncontains"):]
    checked, problems = 0, []
    for sub in re.split(r"\n\s*(?=subroutine\s+\w+)", body, flags=re.I):
        env = {}
        for m in re.finditer(r"(?:type|class)\s*\(\s*(\w+)\s*\)[^:\n]*::\s*([^\n]+)", sub, flags=re.I):
            for n in re.split(r",(?![^()]*\))", m.group(2)):
                env[re.sub(r"\(.*|=.*", "", n).strip().lower()] = m.group(1).lower()

        def resolve(names):
            t = env.get(names[0])
            if t is None:
                return "?"                          # not a derived-type variable of this scope
            for c in names[1:]:
                if t is None:
                    return "component %s of an intrinsic / procedure component" % c
                if t not in types:
                    return "?"                      # a type defined outside the parsed sources
                if c not in types[t]:
                    return "type %s has no component %s" % (t, c)
                t = types[t][c]
            return t

        # associate aliases, in order of appearance (an alias may build on an earlier one)
        for m in re.finditer(r"associate\s*\((.*?)\)\s*\n", sub, flags=re.I | re.S):
            for part in re.split(r",(?![^()]*\))", m.group(1)):
                if "=>" not in part:
                    continue
                alias, expr = part.split("=>", 1)
                ch = _chains(expr)
                tgt = ch[0] if ch else [expr.strip().lower()]
                r = resolve(tgt) if len(tgt) > 1 else env.get(tgt[0])
                env[alias.strip().lower()] = r if (r in types) else env.get(alias.strip().lower())
                if isinstance(r, str) and r.startswith(("type ", "component ")):
                    problems.append("%s => %s: %s" % (alias.strip(), expr.strip(), r))
        # actual arguments of the bound C functions must be intrinsic data, never a derived type
        # (e.g. a growing array must be passed as its %val)
        for m in re.finditer(r"\b(musb200_\w+)\s*\(", sub):
            k, depth = m.end(), 1
            while k < len(sub) and depth:
                depth += sub[k] == "("
                depth -= sub[k] == ")"
                k += 1
            for arg in re.split(r",(?![^()]*\))", sub[m.end():k - 1]):
                ch = _chains(arg)
                pure = len(ch) == 1 and re.fullmatch(r"\s*[A-Za-z_]\w*(\s*\([^()]*\))?(\s*%\s*\w+(\s*\([^()]*\))?)+\s*", arg)
                if pure:
                    r = resolve(ch[0])
                    if r in types:
                        problems.append("%s(... %s ...): a %s is passed, not data" % (m.group(1), arg.strip(), r))
        for names in _chains(sub):
            r = resolve(names)
            if r != "?":
                checked += 1
            if isinstance(r, str) and r.startswith(("type ", "component ")):
                problems.append("%s: %s" % ("%".join(names), r))
    assert checked > 60, checked
    assert not problems, "\n".join(sorted(set(problems)))
