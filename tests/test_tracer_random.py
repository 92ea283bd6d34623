"""Randomised checks of the callable-lowering chain (trace -> simplify -> differentiate -> emit), CPU only.

Random expression trees over the operations user drift / noise functions are made of are built twice: once through
the hash-consing, simplifying `Expr` constructors and once as plain Python closures on floats.  Checked:
* the simplifier never changes the value (`Expr` evaluation = direct evaluation),
* symbolic derivatives agree with central finite differences of the direct evaluation,
* the emitted straight-line C body (the same text NVRTC compiles, `lang="c"`), built with gcc, returns the same
  values and derivatives for a whole batch of expressions in one translation unit - common subexpressions, `sincos`
  pairing and literal formatting included.
"""
import ctypes
import math
import os
import subprocess
import tempfile

import numpy as np

from stochastic_b200.trace import emit, expr as E

NVARS = 3
UNARY = [
    ("sin", E.Expr.sin, math.sin), ("cos", E.Expr.cos, math.cos), ("tanh", E.Expr.tanh, math.tanh),
    ("exp", lambda a: (a.tanh() * 1.5).exp(), lambda a: math.exp(1.5 * math.tanh(a))),   # bounded argument
    ("atan", E.Expr.arctan, math.atan),
    ("sqrt1p", lambda a: (a * a + 1.0).sqrt(), lambda a: math.sqrt(a * a + 1.0)),
    ("log1psq", lambda a: (a * a + 1.5).log(), lambda a: math.log(a * a + 1.5)),
    ("neg", lambda a: -a, lambda a: -a),
    ("square", lambda a: (a * 0.5) ** 2, lambda a: (0.5 * a) ** 2),
    ("cube", lambda a: (a * 0.5) ** 3, lambda a: (0.5 * a) ** 3),
    ("recip", lambda a: 1.0 / (a * a + 2.0), lambda a: 1.0 / (a * a + 2.0)),
]
BINARY = [
    ("add", lambda a, b: a + b, lambda a, b: a + b), ("sub", lambda a, b: a - b, lambda a, b: a - b),
    ("mul", lambda a, b: a * b, lambda a, b: a * b),
    ("div", lambda a, b: a / (b * b + 1.0), lambda a, b: a / (b * b + 1.0)),
    ("rsub", lambda a, b: 2.0 - a * b, lambda a, b: 2.0 - a * b),
]
CONSTS = [0.0, 1.0, -1.0, 2.0, 0.5, -0.75, 3.0, 1e-3, 1.23456]


def random_tree(rng, depth):
    """Returns (Expr, python closure on a list of floats)."""
    if depth == 0 or rng.random() < 0.15:
        if rng.random() < 0.7:
            i = int(rng.integers(NVARS))
            return E.var(i), (lambda x, i=i: x[i])
        c = float(rng.choice(CONSTS))
        return E.const(c), (lambda x, c=c: c)
    if rng.random() < 0.45:
        _, fe, fp = UNARY[int(rng.integers(len(UNARY)))]
        ea, pa = random_tree(rng, depth - 1)
        return E.wrap(fe(ea)), (lambda x, fp=fp, pa=pa: fp(pa(x)))
    _, fe, fp = BINARY[int(rng.integers(len(BINARY)))]
    ea, pa = random_tree(rng, depth - 1)
    eb, pb = random_tree(rng, depth - 1)
    return E.wrap(fe(ea, eb)), (lambda x, fp=fp, pa=pa, pb=pb: fp(pa(x), pb(x)))


def _cases(seed, n, depth):
    """n random trees whose values stay moderate on the sampling box (nested products can grow without bound)."""
    rng = np.random.default_rng(seed)
    probe = np.random.default_rng(seed + 1000).uniform(-1.5, 1.5, size=(16, NVARS))
    out = []
    while len(out) < n:
        e, f = random_tree(rng, depth)
        try:
            vals = [f(list(p)) for p in probe]
        except (OverflowError, ZeroDivisionError, ValueError):
            continue
        if all(math.isfinite(v) and abs(v) < 1e6 for v in vals):
            out.append((e, f))
    return out, rng


def test_simplifier_preserves_values():
    cases, rng = _cases(11, 300, 5)
    pts = rng.uniform(-1.5, 1.5, size=(7, NVARS))
    for e, f in cases:
        got = E.evaluate([e], pts)[0]
        want = np.array([f(list(p)) for p in pts])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12), repr(e)[:200]


def test_symbolic_derivatives_match_finite_differences():
    cases, rng = _cases(12, 200, 4)
    pts = rng.uniform(-1.2, 1.2, size=(3, NVARS))
    h = 1e-5
    for e, f in cases:
        for v in range(NVARS):
            de = E.diff(e, v)
            got = E.evaluate([de], pts)[0]
            for k, p in enumerate(pts):
                xp, xm = list(p), list(p)
                xp[v] += h
                xm[v] -= h
                fd = (f(xp) - f(xm)) / (2 * h)
                scale = max(1.0, abs(fd))
                assert abs(got[k] - fd) < 2e-6 * scale + 1e-7 * max(1.0, abs(f(list(p)))), (repr(e)[:200], v)
            # second derivatives are what the Wagner-Platen tensors need: d/dv of d/dv exists and is finite here
            d2 = E.evaluate([E.diff(de, v)], pts)[0]
            assert np.all(np.isfinite(d2))


def test_emitted_c_matches_interpreter_for_a_batch_of_expressions():
    cases, rng = _cases(13, 60, 5)
    outs, exprs = [], []
    for k, (e, _) in enumerate(cases):
        exprs.append(e)
        outs.append((f"out[{len(outs)}]", e))
        outs.append((f"out[{len(outs)}]", E.diff(e, k % NVARS)))
        exprs.append(E.diff(e, k % NVARS))
    names = {i: f"x[{i}]" for i in range(NVARS)}
    body = emit.emit_body(outs, names, f64=True, lang="c")
    src = ("#include <math.h>\ntypedef double real;\n"
           "static inline void sincos_(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }\n"
           "void f(const real* x, real* out) {\n" + body + "\n}\n")
    tmp = tempfile.mkdtemp()
    c, so = os.path.join(tmp, "f.c"), os.path.join(tmp, "f.so")
    with open(c, "w") as fh:
        fh.write(src)
    subprocess.check_call(["gcc", "-O1", "-D_GNU_SOURCE", "-shared", "-fPIC", "-o", so, c, "-lm"])
    lib = ctypes.CDLL(so)
    lib.f.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    for p in rng.uniform(-1.3, 1.3, size=(5, NVARS)):
        x = np.ascontiguousarray(p)
        out = np.zeros(len(outs))
        lib.f(x.ctypes.data, out.ctypes.data)
        want = np.array([float(v) for v in E.evaluate(exprs, x)])
        assert np.allclose(out, want, rtol=1e-12, atol=1e-12)


def test_reciprocal_division_option_keeps_values_to_the_last_bits():
    """config "reciprocal_division": shared denominators are inverted once.  Same values up to a couple of ulp, fewer
    divisions in the text, and the default (off) text is unchanged."""
    from stochastic_b200 import workloads
    from stochastic_b200.trace.taylor import TracedProblem, step_expressions
    prob = workloads.polar_walk(4.0, (2.0, 0.0), f64=True)
    tp = TracedProblem(prob.a, prob.b, 2)
    outs, sv = step_expressions(tp, "wagner_platen")
    names = {i: f"x[{i}]" for i in range(sv.count)}
    texts = {}
    libs = {}
    for flag in (False, True):
        body = emit.emit_body([(f"out[{i}]", e) for i, e in enumerate(outs)], names, f64=True, lang="c",
                              reciprocal_division=flag)
        texts[flag] = body
        src = ("#define _GNU_SOURCE\n#include <math.h>\n#include <stdbool.h>\ntypedef double real;\n"
               "void f(const real* x, real* out) {\n" + body + "\n}\n")
        tmp = tempfile.mkdtemp()
        c, so = os.path.join(tmp, "f.c"), os.path.join(tmp, "f.so")
        with open(c, "w") as fh:
            fh.write(src)
        subprocess.check_call(["gcc", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, c, "-lm"])
        libs[flag] = ctypes.CDLL(so)
        libs[flag].f.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert texts[False].count(" / ") >= 10 and texts[True].count(" / ") <= 3
    assert emit.emit_body([(f"out[{i}]", e) for i, e in enumerate(outs)], names, f64=True, lang="c") == texts[False]
    rng = np.random.default_rng(3)
    for _ in range(20):
        x = np.concatenate([[rng.uniform(0.5, 3.0), rng.uniform(-3, 3)], rng.normal(size=sv.count - 2) * 0.1])
        x[sv.dt] = 2.0 ** -5
        o0, o1 = np.zeros(len(outs)), np.zeros(len(outs))
        libs[False].f(x.ctypes.data, o0.ctypes.data)
        libs[True].f(x.ctypes.data, o1.ctypes.data)
        assert np.allclose(o0, o1, rtol=1e-14, atol=1e-15)
