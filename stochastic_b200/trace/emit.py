"""DAG -> CUDA C++ device-function text.

The emitted text is straight-line SSA (`const real tN = ...;`) over the hash-consed DAG, so
every shared subexpression (a(x), b(x), their derivatives, sin/cos pairs) is computed once.
It is compiled for sm_100a by NVRTC at run time (or nvcc at build time) together with
`csrc/sde_core.cuh` / `csrc/sde_kernel.cuh`.  With `lang="c"` the same body is emitted as plain
C (math.h) so that the tests can run the generated code on the host against the oracle.
"""
import numpy as np

from . import expr as E

_FUNC = {
    "sin": "sin", "cos": "cos", "tan": "tan", "exp": "exp", "log": "log", "sqrt": "sqrt", "tanh": "tanh",
    "sinh": "sinh", "cosh": "cosh", "atan": "atan", "asin": "asin", "acos": "acos", "abs": "fabs",
    "log1p": "log1p", "expm1": "expm1", "floor": "floor", "trunc": "trunc", "erf": "erf",
    "pow": "pow", "max": "fmax", "min": "fmin", "fmod": "fmod", "atan2": "atan2",
}
_CMP = {"lt": "<", "le": "<=", "gt": ">", "ge": ">=", "eq": "==", "ne": "!="}


def _literal(v, f64):
    if isinstance(v, bool):
        return "true" if v else "false"
    if f64:
        s = repr(float(v))
    else:
        s = repr(float(np.float32(v)))
    if s in ("inf", "-inf", "nan"):
        base = {"inf": "(1.0/0.0)", "-inf": "(-1.0/0.0)", "nan": "(0.0/0.0)"}[s]
        return base if f64 else f"((float){base})"
    if "e" not in s and "." not in s:
        s += ".0"
    return s if f64 else s + "f"


def emit_body(outputs, var_names, f64=True, lang="cuda", indent="    ", reciprocal_division=None):
    """outputs: list of (lhs_text, Expr).  Returns the statement list as text.

    All temporaries are computed before any output is assigned, so outputs may alias inputs.
    reciprocal_division (default: config "reciprocal_division"): denominators shared by two or more divisions are
    inverted once and multiplied with."""
    if reciprocal_division is None:
        from .. import config
        reciprocal_division = config.get("reciprocal_division")
    roots = [e for _, e in outputs]
    order = E.topo_order(roots)
    den_uses = {}
    if reciprocal_division:
        for n in order:
            if n.op == "div" and n.args[1].op != "const":
                den_uses[n.args[1].id] = den_uses.get(n.args[1].id, 0) + 1
    recip = {}                                   # denominator id -> name of its reciprocal
    # sin/cos of the same argument -> one sincos call
    sin_of, cos_of = {}, {}
    for n in order:
        if n.op == "sin":
            sin_of[n.args[0].id] = n
        elif n.op == "cos":
            cos_of[n.args[0].id] = n
    paired = {aid for aid in sin_of if aid in cos_of}
    name = {}
    lines = []
    suffix = "" if (f64 or lang == "cuda") else "f"   # C needs sinf..., CUDA overloads resolve
    done_pairs = set()

    def ref(n):
        if n.op == "const":
            return _literal(n.value, f64)
        if n.op == "var":
            return var_names[n.value]
        return name[n.id]

    for n in order:
        if n.op in ("const", "var"):
            continue
        t = f"t{len(name)}"
        a = [ref(x) for x in n.args]
        op = n.op
        typ = "bool" if n.is_bool else "real"
        if op in ("sin", "cos") and n.args[0].id in paired:
            aid = n.args[0].id
            if aid not in done_pairs:
                done_pairs.add(aid)
                ts, tc = f"t{len(name)}", f"t{len(name) + 1}"
                name[sin_of[aid].id] = ts
                name[cos_of[aid].id] = tc
                lines.append(f"real {ts}, {tc};")
                lines.append(f"sincos{suffix}({a[0]}, &{ts}, &{tc});")
            continue
        if op == "add":
            rhs = f"{a[0]} + {a[1]}"
        elif op == "sub":
            rhs = f"{a[0]} - {a[1]}"
        elif op == "mul":
            rhs = f"{a[0]} * {a[1]}"
        elif op == "div":
            den = n.args[1]
            if den_uses.get(den.id, 0) >= 2:
                if den.id not in recip:
                    recip[den.id] = f"r{len(recip)}"
                    lines.append(f"const real {recip[den.id]} = (real)1 / {a[1]};")
                if n.args[0].op == "const" and n.args[0].value == 1.0:
                    name[n.id] = recip[den.id]   # 1 / den is the reciprocal itself
                    continue
                rhs = f"{a[0]} * {recip[den.id]}"
            else:
                rhs = f"{a[0]} / {a[1]}"
        elif op == "neg":
            rhs = f"-{a[0]}"
        elif op == "sign":
            rhs = f"(real)(({a[0]} > 0) - ({a[0]} < 0))"
        elif op == "rsqrt":
            # CUDA: the math library's rsqrt (1 ulp in fp64, no division); fp32 and host C: 1 / sqrt, both correctly rounded
            rhs = f"rsqrt({a[0]})" if (lang == "cuda" and f64) else f"(real)1 / sqrt{suffix}({a[0]})"
        elif op == "where":
            rhs = f"{a[0]} ? {a[1]} : {a[2]}"
        elif op in _CMP:
            rhs = f"{a[0]} {_CMP[op]} {a[1]}"
        elif op == "and":
            rhs = f"{a[0]} && {a[1]}"
        elif op == "or":
            rhs = f"{a[0]} || {a[1]}"
        elif op == "not":
            rhs = f"!{a[0]}"
        elif op in _FUNC:
            rhs = f"{_FUNC[op]}{suffix}({', '.join(a)})"
        else:
            raise NotImplementedError(f"cannot emit op {op}")
        name[n.id] = t
        lines.append(f"const {typ} {t} = {rhs};")
    for lhs, e in outputs:
        lines.append(f"const real o_{len(lines)} = {ref(e)};  // -> {lhs}")
    # assign after everything has been read
    assigns = []
    k = 0
    for ln in lines:
        if ln.startswith("const real o_"):
            var = ln.split()[2]
            assigns.append(f"{outputs[k][0]} = {var};")
            k += 1
    return "\n".join(indent + ln for ln in lines + assigns)


def count_ops(roots):
    """Operation census of a DAG (for DESIGN.md / diagnostics)."""
    census = {}
    for n in E.topo_order(roots):
        if n.op not in ("const", "var"):
            census[n.op] = census.get(n.op, 0) + 1
    return census
