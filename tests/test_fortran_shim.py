"""The Fortran side of the boundary (fortran/mwgpu_mod.f90) cannot be compiled in this image (no Fortran compiler), so
its agreement with the C ABI is checked statically: every bind(C) derived type must list the members of its C struct
in the same order with interoperable kinds, and every bind(C) interface must name a function include/mwgpu.h declares,
with the same number of arguments, `value` exactly where C passes by value and the matching kind for each scalar."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "mwgpu.h")).read()
F90 = open(os.path.join(ROOT, "fortran", "mwgpu_mod.f90")).read()

KIND = {"double": "real(c_double)", "int": "integer(c_int)", "int64_t": "integer(c_int64_t)", "int32_t": "integer(c_int32_t)",
        "float": "real(c_float)", "char": "character(kind=c_char)", "unsigned char": "character(kind=c_char)"}


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def c_structs():
    out = {}
    for m in re.finditer(r"typedef struct (\w+)\s*\{(.*?)\}\s*\1\s*;", _strip_c_comments(HDR), flags=re.S):
        members = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, names = re.match(r"((?:unsigned\s+)?\w+)\s+(.*)", decl, flags=re.S).groups()
            for nm in names.split(","):
                a = re.match(r"\s*(\w+)\s*(?:\[(\d+)\])?\s*$", nm)
                members.append((KIND[ty], a.group(1).lower(), int(a.group(2) or 1)))
        out[m.group(1)] = members
    return out


def _join_continuations(s):
    s = re.sub(r"!.*", "", s)                      # comments (no '!' inside strings in the interface part)
    return re.sub(r"&\s*\n\s*", " ", s)


def f_types():
    out = {}
    for m in re.finditer(r"type\s*,\s*bind\(C\)\s*::\s*(\w+)(.*?)end type", _join_continuations(F90), flags=re.S | re.I):
        members = []
        for line in m.group(2).strip().splitlines():
            line = line.strip()
            if not line:
                continue
            ty, names = [x.strip() for x in line.split("::")]
            for nm in names.split(","):
                a = re.match(r"\s*(\w+)\s*(?:\((\d+)\))?\s*$", nm)
                members.append((re.sub(r"\s+", "", ty), a.group(1).lower(), int(a.group(2) or 1)))
        out[m.group(1)] = members
    return out


def c_functions():
    out = {}
    for m in re.finditer(r"^\s*(?:const\s+)?(\w+)\s*\*?\s*(mwgpu_\w+)\s*\(([^;{]*?)\)\s*;", _strip_c_comments(HDR), flags=re.M | re.S):
        args = m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                ptr = a.count("*")
                base = re.sub(r"\bconst\b|\*", " ", a).split()
                params.append((ptr, " ".join(base[:-1])))          # (pointer depth, C type)
        out[m.group(2)] = params
    return out


def f_interfaces():
    src = _join_continuations(F90)
    out = {}
    pat = r"(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C\s*,\s*name='(\w+)'\)(.*?)end (?:function|subroutine)"
    for m in re.finditer(pat, src, flags=re.S | re.I):
        dummies = [d.strip().lower() for d in m.group(2).split(",") if d.strip()]
        decl = {}
        for line in m.group(4).splitlines():
            if "::" not in line or line.strip().lower().startswith("import"):
                continue
            attrs, names = line.split("::")
            attrs = re.sub(r"\s+", "", attrs).lower()
            for nm in re.split(r",(?![^()]*\))", names):
                nm = re.match(r"\s*(\w+)", nm).group(1).lower()
                decl[nm] = attrs
        out[m.group(3)] = (m.group(1), dummies, decl)
    return out


def test_the_parsers_see_what_is_there():
    cs, ft, cf, fi = c_structs(), f_types(), c_functions(), f_interfaces()
    assert {"mwgpu_mc_params", "mwgpu_walker_state", "mwgpu_therm_row", "mwgpu_flat_params", "mwgpu_flat_report"} <= set(cs)
    assert len(cf) >= 50 and len(fi) >= 35
    assert "mwgpu_mc_run" in cf and "mwgpu_mc_run" in fi


@pytest.mark.parametrize("name", sorted(f_types()))
def test_derived_types_mirror_the_c_structs(name):
    cs = c_structs()
    assert name in cs, f"{name}: no such struct in include/mwgpu.h"
    assert f_types()[name] == cs[name]


@pytest.mark.parametrize("cname", sorted(f_interfaces()))
def test_interfaces_match_the_c_prototypes(cname):
    cf = c_functions()
    assert cname in cf, f"{cname}: bound by the Fortran module, not declared in include/mwgpu.h"
    fname, dummies, decl = f_interfaces()[cname]
    assert fname == cname                                   # the module keeps the C names
    params = cf[cname]
    assert len(dummies) == len(params), f"{cname}: {len(dummies)} Fortran dummies, {len(params)} C parameters"
    for d, (ptr, cty) in zip(dummies, params):
        assert d in decl, f"{cname}: dummy {d} is not declared"
        attrs = decl[d]
        by_value = ",value" in attrs
        if cty == "mwgpu_ctx":
            # the context is an opaque handle: mwgpu_ctx* = type(c_ptr),value ; mwgpu_ctx** (out) = type(c_ptr)
            assert attrs.startswith("type(c_ptr)"), f"{cname}({d}): the context is an opaque pointer"
            assert by_value == (ptr == 1), f"{cname}({d}): C has {ptr} level(s) of indirection, Fortran says {attrs}"
            continue
        assert by_value == (ptr == 0), f"{cname}({d}): C passes {'a pointer' if ptr else 'by value'}, Fortran says {attrs}"
        if cty in KIND:
            assert attrs.startswith(KIND[cty].lower()), f"{cname}({d}): C {cty}, Fortran {attrs}"
        elif cty.startswith("mwgpu_"):
            assert attrs.startswith(f"type({cty})"), f"{cname}({d}): C {cty}, Fortran {attrs}"


def _call_sites(path):
    """(function, number of actual arguments, line) of every mwgpu_* reference in a Fortran source file."""
    src = open(path).read()
    lines = []
    for ln in src.splitlines():
        out, q = [], None
        for ch in ln:                                   # strip comments, keeping '!' inside character literals
            if q:
                q = None if ch == q else q
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                break
            out.append(ch)
        lines.append("".join(out))
    text = re.sub(r"&\s*\n\s*&?", " ", "\n".join(lines))
    sites = []
    for m in re.finditer(r"\b(mwgpu_\w+)\s*\(", text):
        depth, i, args, cur, q = 1, m.end(), [], [], None
        while depth and i < len(text):
            ch = text[i]
            if q:
                q = None if ch == q else q
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    break
            elif ch == "," and depth == 1:
                args.append("".join(cur)); cur = []; i += 1
                continue
            cur.append(ch); i += 1
        if "".join(cur).strip() or args:
            args.append("".join(cur))
        sites.append((m.group(1), len(args), text.count("\n", 0, m.start()) + 1))
    return sites


@pytest.mark.parametrize("fname", ["molint_gpu.F90", "mc_cycle_gpu.F90"])
def test_call_sites_of_the_host_fragments_match_the_interfaces(fname):
    """Every libmwgpu call in the drop-in `module energy` and in the mc_cycle fragment names an interface of the
    module and passes as many arguments as it declares."""
    fi = f_interfaces()
    derived = set(f_types())
    helpers = {"mwgpu_check": 2}
    sites = _call_sites(os.path.join(ROOT, "fortran", fname))
    assert len(sites) >= 10
    for name, nargs, line in sites:
        if name in derived:                             # type(mwgpu_xxx) :: declarations
            continue
        if name in helpers:
            assert nargs == helpers[name], f"{fname}:{line}: {name} takes {helpers[name]} arguments"
            continue
        assert name in fi, f"{fname}:{line}: {name} has no interface in mwgpu_mod.f90"
        assert nargs == len(fi[name][1]), f"{fname}:{line}: {name} called with {nargs} arguments, interface has {len(fi[name][1])}"
