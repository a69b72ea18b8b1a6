#!/usr/bin/env python3
"""Public API of a concrete-ntt-shaped Rust source tree: {module: {Type or "<module>": {fn: arity}}} for the inherent methods of
the public plan / divisor types and the module-level `pub fn`s of the helper modules.  Used to diff the Rust facade (rust/src)
against the reference (src/ of zama-ai/concrete-ntt): tests/test_rust_facade_api.py.   usage: extract_rust_api.py SRC_DIR"""
import json
import os
import re
import sys

TYPES = ("Plan", "Plan32", "Plan52", "Div32", "Div64", "Solinas")
MODULES = ("prime32", "prime64", "native32", "native64", "native128", "native_binary32", "native_binary64", "native_binary128",
           "product", "fastdiv", "prime")


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def arity(params):
    """number of parameters except self, commas inside <>, (), [] ignored"""
    depth, parts, cur = 0, [], ""
    for ch in params:
        if ch in "<([":
            depth += 1
        elif ch in ">)]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    parts = [p.strip() for p in parts if p.strip()]
    return len([p for p in parts if not re.match(r"^(&\s*)?(mut\s+)?self\b", p)])


def fns_in(body):
    out = {}
    for m in re.finditer(r"\bpub\s+(?:const\s+)?(?:unsafe\s+)?fn\s+(\w+)\s*(?:<[^>]*>)?\s*\(", body):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(body[i], 0)
            i += 1
        out[m.group(1)] = arity(body[m.end():i - 1])
    return out


def block_after(src, start):
    """text of the {...} block whose opening brace is the first one at or after `start`"""
    i = src.index("{", start)
    depth, j = 1, i + 1
    while depth:
        depth += {"{": 1, "}": -1}.get(src[j], 0)
        j += 1
    return src[i + 1:j - 1], j


def expand_macros(src):
    """the facade generates its native modules with macro_rules: read the macro arms as plain text, so that `impl Plan32 { .. }`
    inside `macro_rules! plan32` is attributed to every module that invokes it"""
    return src


def api_of_file(src):
    src = strip_comments(src)
    # drop #[cfg(test)] modules
    for m in list(re.finditer(r"#\[cfg\(test\)\]\s*(?:pub\s+)?mod\s+\w+\s*", src))[::-1]:
        try:
            _, end = block_after(src, m.end() - 1)
            src = src[:m.start()] + src[end:]
        except ValueError:
            pass
    api = {}
    for m in re.finditer(r"^\s*impl\s+(\w+)\s*\{", src, flags=re.M):
        if m.group(1) in TYPES:
            body, _ = block_after(src, m.end() - 1)
            api.setdefault(m.group(1), {}).update(fns_in(body))
    # module-level pub fns: text outside every impl / mod / trait block is approximated by column-0 declarations
    top = {}
    for m in re.finditer(r"^pub\s+(?:const\s+)?fn\s+(\w+)\s*(?:<[^>]*>)?\s*\(", src, flags=re.M):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        top[m.group(1)] = arity(src[m.end():i - 1])
    if top:
        api["<module>"] = top
    return api


def reference_api(src_dir):
    out = {}
    for mod in MODULES:
        path = os.path.join(src_dir, mod + ".rs")
        if os.path.exists(path):
            out[mod] = api_of_file(open(path).read())
    return out


if __name__ == "__main__":
    print(json.dumps(reference_api(sys.argv[1]), indent=1, sort_keys=True))


def facade_api(src_dir):
    """the same structure for the Rust facade (rust/src): the native modules live in one file and `prime` inside lib.rs"""
    out = {}
    for mod in ("prime32", "prime64", "product", "fastdiv"):
        out[mod] = api_of_file(open(os.path.join(src_dir, mod + ".rs")).read())
    for fname in ("native.rs", "lib.rs"):
        src = strip_comments(open(os.path.join(src_dir, fname)).read())
        for m in re.finditer(r"^pub\s+mod\s+(\w+)\s*\{", src, flags=re.M):
            if m.group(1) in MODULES:
                body, _ = block_after(src, m.end() - 1)
                # module-level fns of an inline module are indented: dedent one level before the column-0 scan
                body = "\n".join(line[4:] if line.startswith("    ") else line for line in body.split("\n"))
                out[m.group(1)] = api_of_file(body)
    return out
