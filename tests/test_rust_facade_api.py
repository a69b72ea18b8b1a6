"""The Rust facade (rust/src, not compilable here: no cargo) exposes every public method of the reference's plan types and helper
modules under the same module path, type name, method name and number of arguments.  The reference side of the comparison is
tests/golden/reference_api_v1.json, extracted from the crate's sources by tools/extract_rust_api.py and re-checked against
/root/reference whenever that tree is present."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import extract_rust_api as X  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "reference_api_v1.json")))


def test_reference_api_fixture_is_current():
    src = "/root/reference/src"
    if os.path.isdir(src):
        assert json.loads(json.dumps(X.reference_api(src))) == GOLD


def test_fixture_covers_the_north_star_surface():
    for mod in ("prime32", "prime64", "native32", "native64", "native128", "native_binary32", "native_binary64", "native_binary128", "product"):
        assert mod in GOLD
    assert set(GOLD["prime32"]["Plan"]) == {"try_new", "ntt_size", "modulus", "fwd", "inv", "mul_assign_normalize", "normalize", "mul_accumulate"}
    assert GOLD["native64"]["Plan32"]["fwd"] == 6 and GOLD["native128"]["Plan32"]["inv"] == 11


def test_facade_has_every_reference_method_with_the_same_arity():
    facade = X.facade_api(os.path.join(ROOT, "rust", "src"))
    missing = []
    for mod, types in GOLD.items():
        for ty, fns in types.items():
            have = facade.get(mod, {}).get(ty)
            if have is None:
                missing.append("%s::%s" % (mod, ty))
                continue
            for fn, ar in fns.items():
                if have.get(fn) != ar:
                    missing.append("%s::%s::%s/%d (facade: %r)" % (mod, ty, fn, ar, have.get(fn)))
    assert not missing, missing


def test_native_rs_is_the_generator_output():
    import subprocess
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "gen_rust_native.py")]).decode()
    assert out == open(os.path.join(ROOT, "rust", "src", "native.rs")).read()
