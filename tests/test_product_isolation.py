"""The product must not route through the checker: nothing under jvector_b200/ or include/ may import, link, load or call
anything under oracle/ (or the reference tree), and the built library must not depend on the oracle's shared objects."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FORBIDDEN = re.compile(r"jv_oracle|libjv_oracle|oracle_lib|oracle/|jvo_|/root/reference|_ref/libjvector")


def _product_files():
    for base in ("jvector_b200", "include"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            if "lib" in dp.split(os.sep)[-2:] and dp.endswith(("lib", "obj")):
                continue
            for f in fn:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    yield os.path.join(dp, f)


def test_product_sources_never_touch_the_oracle():
    offenders = []
    for path in _product_files():
        for ln, line in enumerate(open(path, errors="replace"), 1):
            code = line.split("//")[0] if path.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp")) else line.split("#")[0]
            if FORBIDDEN.search(code) and "reference" not in code.lower().split("/root/reference")[0][-0:]:
                offenders.append("%s:%d: %s" % (os.path.relpath(path, ROOT), ln, line.strip()))
    # comments may CITE the reference (file:line); code may not open it
    offenders = [o for o in offenders if "/root/reference" not in o or "open(" in o or "CDLL" in o]
    assert not offenders, offenders


def test_library_has_no_oracle_dependency():
    from jvector_b200 import _native as nat
    if not os.path.exists(nat.SO):
        from jvector_b200 import build
        build.build()
    needed = subprocess.check_output(["readelf", "-d", nat.SO], text=True)
    assert "oracle" not in needed and "libjvector.so" not in needed
    syms = subprocess.check_output(["nm", "-D", nat.SO], text=True)
    assert "jvo_" not in syms
