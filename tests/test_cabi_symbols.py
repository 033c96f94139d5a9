"""CPU: the C-ABI library builds, loads without a GPU, and exports every symbol the header declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "fasta_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb200_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__
    __graft_entry__.build()
    from fasta import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/fasta_b200.h but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes signature in fasta/_cabi.py"
    assert sorted(_cabi.SIGNATURES) == names
    loaded = _cabi.load()
    assert loaded.fb200_abi_version() == _cabi.ABI_VERSION
    assert loaded.fb200_workspace_bytes(40000, 100000) > 32 * 100096 * 8


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import pytest
    import torch
    import fasta
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fasta._cabi.Fb200Error):
        fasta.fasta(fasta.linalg.LinearMap.identity((3,)), lambda z: 0, lambda z: z, None, None, np.zeros(3),
                    verbose=False)
    with pytest.raises(fasta._cabi.Fb200Error):
        fasta.proximal.shrink(np.ones(4), 0.5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fasta-python_b200", "fasta")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{fn} mentions the oracle"


def test_backend_constructors_assign_before_use():
    """Static check (the GPU back-ends cannot be constructed without a device): no attribute of `self` is read in an
    __init__ of fasta/_backends.py before it has been assigned there."""
    import ast
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fasta-python_b200", "fasta", "_backends.py")
    tree = ast.parse(open(path).read())
    for cls in (n for n in tree.body if isinstance(n, ast.ClassDef)):
        inits = [f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name == "__init__"]
        if not inits:
            continue
        names = {f.name for f in cls.body if isinstance(f, ast.FunctionDef)}
        names |= {t.id for n in cls.body if isinstance(n, ast.Assign) for t in n.targets if isinstance(t, ast.Name)}
        assigned = set()
        for stmt in inits[0].body:
            for n in ast.walk(stmt):
                if (isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "self"
                        and isinstance(n.ctx, ast.Load)):
                    assert n.attr in assigned or n.attr in names, f"{cls.name}.__init__ reads self.{n.attr} (line {n.lineno}) before assigning it"
            for n in ast.walk(stmt):
                if (isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "self"
                        and isinstance(n.ctx, ast.Store)):
                    assigned.add(n.attr)


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/fasta_b200.h against its ctypes signature in fasta/_cabi.py: same number of
    parameters, same kind (pointer / double / int / int64 / size_t) at every position, same return kind -- an ABI drift
    between the two would otherwise only show up as garbage arguments on the GPU.  Also: the enum constants mirrored
    in _cabi.py carry the header's values."""
    from fasta import _cabi
    text = open(os.path.join(ROOT, "include", "fasta_b200.h")).read()
    nocomment = re.sub(r"/\*.*?\*/", "", text, flags=re.S)

    def kind(ctype_decl):
        d = ctype_decl.strip()
        if "*" in d:
            return "ptr"
        base = re.sub(r"\b(const|unsigned)\b", "", d).split()
        base = base[0] if base else ""
        return {"double": "double", "int": "int", "int64_t": "i64", "size_t": "size", "void": "void", "uint32_t": "u32"}[base]

    of_ctype = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_double: "double", ctypes.c_int: "int",
                ctypes.c_int64: "i64", ctypes.c_size_t: "size", ctypes.c_uint32: "u32"}
    protos = re.findall(r"(?m)^\s*((?:const\s+)?[A-Za-z_0-9]+\s*\**)\s*(fb200_[a-zA-Z0-9_]+)\s*\(([^;{]*?)\)\s*;", nocomment)
    assert len(protos) >= 40
    seen = set()
    for ret, name, params in protos:
        seen.add(name)
        restype, argtypes = _cabi.SIGNATURES[name]
        plist = [p for p in (q.strip() for q in params.replace("\n", " ").split(",")) if p and p != "void"]
        kinds = [kind(re.sub(r"\b[A-Za-z_0-9]+\s*(\[[^\]]*\])?$", "", p) if not p.endswith("*") else p) for p in plist]
        assert len(kinds) == len(argtypes), f"{name}: header has {len(kinds)} parameters, _cabi.py {len(argtypes)}"
        kind_of = lambda a: of_ctype.get(a, "ptr" if hasattr(a, "_type_") and not isinstance(a._type_, str) else None)
        for i, (k, a) in enumerate(zip(kinds, argtypes)):
            assert kind_of(a) == k, f"{name}: parameter {i} is {k} in the header, {kind_of(a)} in _cabi.py"
        assert kind_of(restype) == kind(ret), f"{name}: return type"
    assert seen == set(_cabi.SIGNATURES)
    # enum mirrors
    for cname, pyname in (("FB200_S_F", "S_F"), ("FB200_S_RESTART", "S_RESTART"), ("FB200_S_G1_SQ", "S_G1_SQ"),
                          ("FB200_S_AUX3", "S_AUX3"), ("FB200_S_TAU", "S_TAU"), ("FB200_S_TAU_USED", "S_TAU_USED"),
                          ("FB200_S_SKIP", "S_SKIP"), ("FB200_S_SKIPPED", "S_SKIPPED"), ("FB200_S_FRING", "S_FRING")):
        m = re.search(rf"\b{cname}\s*=\s*(\d+)", nocomment)
        assert m and int(m.group(1)) == getattr(_cabi, pyname), cname
    assert int(re.search(r"#define\s+FB200_NSCAL\s+(\d+)", text).group(1)) == _cabi.NSCAL
    assert int(re.search(r"#define\s+FB200_FRING\s+(\d+)", text).group(1)) == _cabi.FRING


def test_call_sites_pass_the_declared_number_of_arguments():
    """Static check of every `....fb200_xxx(...)` call in the product package: the number of positional arguments is
    the number of parameters of the ctypes signature (call sites with starred arguments are counted as 'at least')."""
    import ast
    from fasta import _cabi
    pkg = os.path.join(ROOT, "fasta-python_b200", "fasta")
    checked = 0
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("fb200_"):
                name = node.func.attr
                assert name in _cabi.SIGNATURES, f"{fn}:{node.lineno}: {name} is not declared"
                want = len(_cabi.SIGNATURES[name][1])
                starred = [a for a in node.args if isinstance(a, ast.Starred)]
                plain = len(node.args) - len(starred)
                assert not node.keywords, f"{fn}:{node.lineno}: keyword arguments in a C call"
                if starred:
                    assert plain <= want, f"{fn}:{node.lineno}: {name} gets more than {want} arguments"
                else:
                    assert plain == want, f"{fn}:{node.lineno}: {name} gets {plain} arguments, the C function takes {want}"
                checked += 1
    assert checked >= 60
