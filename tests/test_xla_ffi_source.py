"""CPU: static checks of the XLA-FFI boundary (csrc/xla_ffi/xla_ffi.cc + jaxrl_b200/jax_ffi.py).  jaxlib is not
installable here, so neither file can run; what can be verified without it:

  1. xla_ffi.cc type-checks (g++ -fsyntax-only) against the REAL C ABI header and a minimal stub of the typed-FFI API
     that static_asserts every handler signature against its Bind() chain (tests/xla_ffi_stub/);
  2. every exported compute entry point of include/mapox_b200.h is called by a handler, and every field of every
     argument struct is assigned;
  3. jax_ffi.py registers exactly the handler symbols the .cc defines, and each of its ffi_call sites passes the
     operand / result / attribute counts (and attribute names) its handler binds; input_output_aliases are in range.
The adapter stays UNVERIFIED against a real jaxlib."""

import ast
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC = os.path.join(ROOT, "jaxrl_b200", "csrc", "xla_ffi", "xla_ffi.cc")
HDR = os.path.join(ROOT, "include", "mapox_b200.h")
PY = os.path.join(ROOT, "jaxrl_b200", "jax_ffi.py")

# not compute entry points: queries, error string, test knobs
NON_COMPUTE = {"mpx_last_error", "mpx_version", "mpx_debug_set", "mpx_debug_set_stamps", "mpx_obs_train_supported",
               "mpx_obs_train_tiles"}
# mpx_muon_step == mpx_muon_step_phased(phases=3): one handler (MpxMuonStep) serves both
SERVED_BY = {"mpx_muon_step": "mpx_muon_step_phased"}


def _handlers():
    src = open(CC).read()
    out = {}
    for m in re.finditer(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),\s*(\w+),(.*?)MPX_TRAITS\);", src, re.S):
        name, impl, chain = m.group(1), m.group(2), m.group(3)
        out[name] = dict(impl=impl, args=len(re.findall(r"\.Arg<", chain)), rets=len(re.findall(r"\.Ret<", chain)),
                         attrs=re.findall(r'\.Attr<[^(]*>\("(\w+)"\)', chain))
    return out, src


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_adapter_type_checks_against_the_c_abi():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "xla_ffi_stub"),
                        "-I", os.path.join(ROOT, "include"), CC], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-4000:]


def test_every_compute_entry_point_has_a_handler_and_every_struct_field_is_set():
    hdr = open(HDR).read()
    handlers, src = _handlers()
    exported = set(re.findall(r"^(?:int|size_t|const char\*)\s+(mpx_\w+)\(", hdr, re.M))
    compute = {s for s in exported if s not in NON_COMPUTE and not s.endswith("_bytes")}
    assert len(compute) >= 44
    for sym in sorted(compute):
        target = SERVED_BY.get(sym, sym)
        assert re.search(r"\b%s\(" % target, src), f"no XLA-FFI handler calls {sym}"
    # workspace-size queries are used by the Python side instead
    py = open(PY).read()
    for sym in sorted(s for s in exported if s.endswith("_bytes")):
        assert sym in py, f"{sym} is not used by jax_ffi.py to size a workspace result"
    for m in re.finditer(r"typedef struct \{(.*?)\}\s*(mpx_\w+_args);", hdr, re.S):
        body, sname = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = re.sub(r"^(const\s+)?[\w ]+?[\s\*]+(?=\w+\s*(,|$))", "", decl)      # drop the type of the first declarator
            fields += [n.strip(" *") for n in names.split(",")]
        assert fields, sname
        assert sname in src, f"{sname} is never filled in by a handler"
        for f in fields:
            assert re.search(r"\ba\.%s\b" % f, src), f"{sname}.{f} is never assigned in xla_ffi.cc"
    assert len(handlers) >= 44


def _ffi_calls(tree):
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "_call":
            yield node


def test_python_side_matches_the_handlers():
    handlers, _ = _handlers()
    tree = ast.parse(open(PY).read())
    targets = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "TARGETS":
            targets = [e.value for e in node.value.elts]
    assert targets is not None and sorted(targets) == sorted(handlers), set(targets) ^ set(handlers)
    seen = set()
    for call in _ffi_calls(tree):
        name = call.args[0].value
        assert name in handlers, name
        h = handlers[name]
        seen.add(name)
        results = call.args[1]
        n_res = len(results.elts) if isinstance(results, ast.Tuple) else 1
        n_ops = len(call.args) - 2
        kw = {k.arg: k.value for k in call.keywords}
        aliases = kw.pop("aliases", None)
        assert n_ops == h["args"], f"{name}: {n_ops} operands passed, handler binds {h['args']}"
        assert n_res == h["rets"], f"{name}: {n_res} results declared, handler binds {h['rets']}"
        assert sorted(kw) == sorted(h["attrs"]), f"{name}: attributes {sorted(kw)} vs {sorted(h['attrs'])}"
        if isinstance(aliases, ast.Dict):
            for k, v in zip(aliases.keys, aliases.values):
                assert 0 <= k.value < h["args"] and 0 <= v.value < h["rets"], f"{name}: alias {k.value}->{v.value} out of range"
    # the hot-path ops the north star names must be wired, not only bound
    for must in ("MpxAttnBlockDecode", "MpxGluFused", "MpxGluBwdFused", "MpxAttnTrainFwd", "MpxAttnTrainBwd", "MpxLinear",
                 "MpxLinearWgrad", "MpxGae", "MpxHlGaussLoss", "MpxHlGaussValue", "MpxPolicyHeadSample", "MpxObsEmbedFused",
                 "MpxValueHeadFused"):
        assert must in seen, f"{must} is registered but never called from jax_ffi.py"


def test_aliases_documented_in_the_adapter_match_python():
    """The '// aliases: operand(i) -> r' comments next to the handlers are the contract for input_output_aliases."""
    handlers, src = _handlers()
    doc = {}
    for m in re.finditer(r"aliases:([^\n]*(?:\n//[^\n]*)?)\n(?:.*?\n)*?static ffi::Error (\w+)Impl", src):
        pairs = re.findall(r"\((\d+)\)\s*->\s*(\d+)", m.group(1))
        doc[m.group(2)] = {int(a): int(b) for a, b in pairs}
    tree = ast.parse(open(PY).read())
    for call in _ffi_calls(tree):
        name = call.args[0].value
        kw = {k.arg: k.value for k in call.keywords}
        if isinstance(kw.get("aliases"), ast.Dict):
            got = {k.value: v.value for k, v in zip(kw["aliases"].keys, kw["aliases"].values)}
            impl = handlers[name]["impl"][:-4]
            assert doc.get(impl) == got, f"{name}: python aliases {got} vs documented {doc.get(impl)}"
