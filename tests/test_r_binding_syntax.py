"""CPU: the R .Call glue (bindings/R/src/Rwrapper_b200.c) compiles against the public headers.  There is no R in the
image, so the R API is declared by stub headers (tests/r_stub/) - this checks syntax, the types of every call into
include/stochqn.h / stochqn_b200.h, and that every registered routine exists with the registered arity."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "bindings", "R", "src", "Rwrapper_b200.c")


def test_r_glue_compiles_against_the_public_headers():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-Werror=int-conversion", "-fsyntax-only", "-DUSE_DOUBLE", "-I" + os.path.join(ROOT, "tests", "r_stub"),
                        "-I" + os.path.join(ROOT, "include"), SRC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_registered_routines_match_their_definitions_and_the_r_call_sites():
    src = open(SRC).read()
    registered = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(r_b200_\w+)", \(DL_FUNC\) &\w+, (\d+)\}', src))
    assert len(registered) == 8
    for name, arity in registered.items():
        m = re.search(r"SEXP %s\(([^)]*)\)" % name, src)
        assert m, name
        assert m.group(1).count("SEXP") == arity, name
    rsrc = open(os.path.join(ROOT, "bindings", "R", "R", "optimizers_free_b200.R")).read()
    for m in re.finditer(r'\.Call\("(r_b200_\w+)"', rsrc):
        name = m.group(1)
        assert name in registered, name
        # count the arguments of this .Call (top-level commas up to the matching parenthesis)
        i = rsrc.index("(", m.start())
        depth, commas, j = 0, 0, i
        while True:
            ch = rsrc[j]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    break
            elif ch == "," and depth == 1:
                commas += 1
            j += 1
        assert commas == registered[name], "%s: .Call passes %d arguments, registered %d" % (name, commas, registered[name])
