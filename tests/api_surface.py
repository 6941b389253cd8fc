"""Public API surface of a C++ header: {(scope, function name, number of parameters)} for the member functions of every
class (public section only) and the free functions of every namespace.  Used by tests/test_api_surface.py to check that the
drop-in headers (include/global_body_planner/*.h) declare what the reference's headers declare for the path; names and
arities only — no reference text is stored."""
import re


def _strip(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    src = re.sub(r"^\s*#[^\n]*(?:\\\n[^\n]*)*", " ", src, flags=re.M)
    return src


def _nargs(params):
    params = params.strip()
    if not params or params == "void":
        return 0
    depth, n = 0, 1
    for ch in params:
        if ch in "<([{":
            depth += 1
        elif ch in ">)]}":
            depth -= 1
        elif ch == "," and depth == 0:
            n += 1
    return n


_DECL = re.compile(r"(?:^|[;{}:])\s*((?:[\w:<>,\*&\s]|\[\]|::)*?[\w>\*&])\s+(operator\s*\S+?|~?\w+)\s*\(([^;{}]*?)\)\s*(const)?\s*(?:override)?\s*(?=[;{])", re.S)
_CTOR = re.compile(r"(?:^|[;{}:])\s*(?:explicit\s+)?(~?\w+)\s*\(([^;{}]*?)\)\s*(?=[;{:])", re.S)
_KEYWORDS = {"if", "for", "while", "switch", "return", "sizeof", "catch", "throw", "else", "do", "new", "delete", "typedef", "using"}


def _blocks(src):
    """yield (kind, name, body) for every top-level class / struct / namespace block, recursing into namespaces"""
    for m in re.finditer(r"\b(class|struct|namespace)\s+(\w+)[^;{]*\{", src):
        depth, i = 1, m.end()
        while i < len(src) and depth:
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        yield m.group(1), m.group(2), src[m.end():i - 1]


def _flatten(body):
    """drop the bodies of inline functions and nested blocks: keep depth-0 text, replace {...} by {}"""
    out, depth = [], 0
    for ch in body:
        if ch == "{":
            if depth == 0:
                out.append("{}")
            depth += 1
        elif ch == "}":
            depth -= 1
        elif depth == 0:
            out.append(ch)
    return "".join(out)


def _public_part(kind, body):
    flat = _flatten(body)
    parts = re.split(r"\b(public|protected|private)\s*:", flat)
    access = "public" if kind == "struct" else "private"
    keep = []
    for p in parts:
        if p in ("public", "protected", "private"):
            access = p
        elif access == "public":
            keep.append(p)
    return ";".join(keep)


def surface(text):
    src = _strip(text)
    out = set()
    for kind, name, body in _blocks(src):
        if kind == "namespace":
            flat = _flatten(body)
            for m in _DECL.finditer(flat):
                fn = m.group(2)
                if fn not in _KEYWORDS and "typedef" not in m.group(1) and "return" not in m.group(1).split():
                    out.add((name, fn, _nargs(m.group(3))))
            continue
        pub = _public_part(kind, body)
        for m in _DECL.finditer(pub):
            fn = m.group(2)
            if fn not in _KEYWORDS and "return" not in m.group(1).split():
                out.add((name, fn, _nargs(m.group(3))))
        for m in _CTOR.finditer(pub):
            if m.group(1).lstrip("~") == name:
                out.add((name, m.group(1), _nargs(m.group(2))))
    return out


def _drop_defaults(params):
    out, depth, skip = [], 0, False
    for ch in params:
        if ch in "<([{":
            depth += 1
        elif ch in ">)]}":
            depth -= 1
        if ch == "," and depth == 0:
            skip = False
        if ch == "=" and depth == 0:
            skip = True
        if not skip:
            out.append(ch)
    return " ".join("".join(out).split())


def declarations(text):
    """[(scope kind, scope, return type, name, parameter list without defaults, const?)] for the functions surface() finds
    (constructors, destructors and operators left out): enough to spell a pointer to each of them."""
    src = _strip(text)
    out = []
    for kind, name, body in _blocks(src):
        part = _flatten(body) if kind == "namespace" else _public_part(kind, body)
        for m in _DECL.finditer(part):
            ret, fn = " ".join(m.group(1).split()), m.group(2)
            words = ret.split()
            if fn in _KEYWORDS or fn.startswith("operator") or "return" in words or "typedef" in words or "friend" in words:
                continue
            static = "static" in words
            ret = " ".join(w for w in words if w not in ("static", "virtual", "inline", "explicit"))
            out.append(("namespace" if kind == "namespace" else "class", name, ret, fn, _drop_defaults(m.group(3)), bool(m.group(4)), static))
    return out


def pointer_checks(text):
    """C++ statements that compile only if every declared function exists with exactly this signature"""
    lines = []
    for i, (kind, scope, ret, fn, params, const, static) in enumerate(declarations(text)):
        if kind == "namespace" or static:
            lines.append(f"{{ using namespace {scope}; {ret} (*p{i})({params}) = &{scope}::{fn}; (void) p{i}; }}" if kind == "namespace"
                         else f"{{ {ret} (*p{i})({params}) = &{scope}::{fn}; (void) p{i}; }}")
        else:
            lines.append(f"{{ {ret} ({scope}::*p{i})({params}){' const' if const else ''} = &{scope}::{fn}; (void) p{i}; }}")
    return lines
