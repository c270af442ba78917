"""The Haskell shim (hs/Numeric/LinearAlgebra/Sparse/B200.hs) cannot be compiled in this image (no GHC), so it is at least kept HONEST
against the C ABI mechanically: every `foreign import ccall` must name a symbol that include/sla_b200.h declares and the library
exports, with the same number of arguments and a compatible shape per argument (pointer vs 32-bit vs 64-bit integer vs double) as the
ctypes table that tests/test_abi.py checks against the header and the .so."""
import ctypes as C
import os
import re

from sparse_linear_algebra_b200 import _lib

HS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hs", "Numeric", "LinearAlgebra", "Sparse", "B200.hs")
IMPORT = re.compile(r'^foreign import ccall (?:safe|unsafe)\s+"(&?)(\w+)"\s+\w+\s*::\s*(.*)$')


def _split_arrows(sig):
    """Top-level split of `a -> b -> IO r` (parentheses may hold arrows: FunPtr (Ptr X -> IO ()))."""
    parts, depth, cur = [], 0, ""
    i = 0
    while i < len(sig):
        ch = sig[i]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if depth == 0 and sig.startswith("->", i):
            parts.append(cur.strip()); cur = ""; i += 2
            continue
        cur += ch
        i += 1
    parts.append(cur.strip())
    return parts


def _shape_hs(t):
    t = t.strip()
    if t.startswith(("Ptr", "FunPtr", "CString")) or t.startswith("(Ptr"):
        return "ptr"
    if t in ("CInt", "Status", "CUInt", "Int32", "Word32"):
        return "i32"
    if t in ("Int64", "Word64", "CLong", "CULong", "CLLong", "CULLong", "CSize"):
        return "i64"
    if t in ("Double", "CDouble"):
        return "f64"
    return "?" + t


def _shape_ct(t):
    if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
        return "ptr"
    if t in (C.c_int, C.c_uint, C.c_int32, C.c_uint32):
        return "i32"
    if t in (C.c_int64, C.c_uint64, C.c_size_t, C.c_longlong, C.c_ulonglong):
        return "i64"
    if t is C.c_double:
        return "f64"
    return "?" + repr(t)


def test_every_foreign_import_matches_the_abi():
    lines = open(HS, encoding="utf-8").read().splitlines()
    # an import may continue on the following (indented) lines
    decls, cur = [], None
    for ln in lines:
        if ln.startswith("foreign import"):
            if cur:
                decls.append(cur)
            cur = ln
        elif cur is not None and ln[:1] in (" ", "\t") and ln.strip():
            cur += " " + ln.strip()
        else:
            if cur:
                decls.append(cur)
            cur = None
    if cur:
        decls.append(cur)
    assert len(decls) >= 40
    L = _lib.load()
    seen = 0
    for d in decls:
        m = IMPORT.match(re.sub(r"\s+--.*$", "", d))
        assert m, f"unparsed foreign import: {d}"
        addr, name, sig = m.groups()
        assert name in _lib.SIGNATURES, f"{name}: imported by the shim, not part of the ABI table"
        assert hasattr(L, name), f"{name}: not exported by the library"
        if addr:                                    # "&sla_x_free": a finalizer pointer, no call signature to compare
            continue
        res, args = _lib.SIGNATURES[name]
        parts = _split_arrows(sig)
        hs_args, hs_res = parts[:-1], parts[-1]
        assert hs_res.startswith("IO"), (name, hs_res)
        assert len(hs_args) == len(args), f"{name}: the shim passes {len(hs_args)} arguments, the ABI takes {len(args)}"
        for k, (h, c) in enumerate(zip(hs_args, args)):
            assert _shape_hs(h) == _shape_ct(c), f"{name}: argument {k}: {h} vs {c}"
        r = hs_res[2:].strip().strip("()").strip()
        if res is None:
            assert r == "", (name, hs_res)
        else:
            assert _shape_hs(r) == _shape_ct(res), f"{name}: result {hs_res} vs {res}"
        seen += 1
    assert seen >= 40
