"""Equation strings -> CUDA translation unit (the GPU counterpart of R/prepareModel.R:479-512, 737-770).

The reference wraps the user's statements into ``latentEquation`` / ``measurementEquation`` C++ functions, binds
every identifier that names an entry of ``parameters`` to an ``arma::mat`` / ``double`` pulled out of an
``Rcpp::List`` by string key on EVERY call, and pastes the result into the filter template.  Here the same
statements are pasted into ``__device__`` members of a model struct; a referenced container becomes a fixed-size
``psy::Mat<r,c>`` (or ``double``) whose free entries read the instance's slot array ``psy_p_[k]`` and whose fixed
entries are literal constants — so structural zeros fold away at compile time and nothing is looked up by name.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class ModelSpec:
    nlatent: int
    nmanifest: int
    containers: list          # model.Container, in parameterList order
    n_slots: int              # free slots = rows of parameterTableInit
    latentEquations: str
    manifestEquations: str
    srUpdate: bool = True
    stepper: int = 0          # 0 = rk4 (also the "default" branch, R/modelTemplate.R:334-344), 1 = runge_kutta_dopri5

    def container(self, name):
        for c in self.containers:
            if c.name == name:
                return c
        raise KeyError(name)


def referenced(equations: str, spec: ModelSpec) -> List[str]:
    """prepareEquations (R/prepareModel.R:737-766): identifiers of the equation text that name a parameter container."""
    toks = [t for t in re.split(r"[^A-Za-z0-9_]", equations) if t]
    names = [c.name for c in spec.containers]
    out = []
    for t in toks:
        if t in names and t not in out:
            out.append(t)
    return out


def lit(v: float) -> str:
    if np.isnan(v):
        return "(0.0/0.0)"
    if np.isinf(v):
        return "(1.0/0.0)" if v > 0 else "(-1.0/0.0)"
    s = repr(float(v))
    if "e" not in s and "." not in s and "n" not in s:
        s += ".0"
    return s


def _elem(c, i, j) -> str:
    k = int(c.slots[i, j])
    return f"psy_p_[{k}]" if k >= 0 else lit(c.values[i, j])


def _bind(c) -> str:
    """Declaration of one container inside an equation function."""
    if c.kind == "double":
        return f"    const double {c.name} = {_elem(c, 0, 0)};\n"
    r, co = c.shape
    lines = [f"    psy::Mat<{r},{co}> {c.name};\n"]
    for j in range(co):
        for i in range(r):
            lines.append(f"    {c.name}.a[{i + j * r}] = {_elem(c, i, j)};\n")
    return "".join(lines)


def _structure(c):
    """(is_diagonal, is_lower) by STRUCTURE: an entry counts as zero only if it is fixed at exactly 0."""
    r, co = c.shape
    diag = lower = True
    for i in range(r):
        for j in range(co):
            if i == j:
                continue
            nz = c.slots[i, j] >= 0 or c.values[i, j] != 0.0
            if nz:
                diag = False
                if j > i:
                    lower = False
    return diag, lower


_MATH_OK = {"exp", "log", "sqrt", "pow", "sin", "cos", "tan", "tanh", "sinh", "cosh", "fabs", "abs", "log1p", "expm1", "atan",
            "std", "t", "person"}


def drift_dependency(spec: ModelSpec) -> List[int]:
    """Sparsity of the drift's Jacobian, as far as it can be read off the equation text: mask[i] has bit r set if
    dx_i MAY depend on x_r.  Conservative — anything that is not a plain `dx(i) = <expression in x(k), parameters, t>;`
    statement (or `dx = MATRIX*x;` with structurally zero entries) makes every component depend on every state.
    The kernel uses it to skip the sigma-point columns that cannot change a component: there f(x+) = f(x-) = f(x0)
    bit for bit, so the difference is exactly 0 and the sum exactly 2 f(x0) in the reference as well."""
    n = spec.nlatent
    full = (1 << n) - 1
    names = {c.name: c for c in spec.containers}
    stmts = [st.strip() for st in spec.latentEquations.replace("\n", " ").split(";") if st.strip()]
    mask = [None] * n
    for st in stmts:
        m = re.match(r"^dx\s*[\(\[]\s*(\d+)\s*[\)\]]\s*=(?!=)\s*(.+)$", st, flags=re.S)
        if m:
            i, expr = int(m.group(1)), m.group(2)
            if i >= n or mask[i] is not None:
                return [full] * n                       # out of range / assigned twice: give up
            deps = 0
            rest = re.sub(r"\bx\s*[\(\[]\s*(\d+)\s*[\)\]]", lambda q: f" @{q.group(1)} ", expr)
            for k in re.findall(r"@(\d+)", rest):
                if int(k) >= n:
                    return [full] * n
                deps |= 1 << int(k)
            idents = set(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", re.sub(r"@\d+", " ", rest)))
            idents = {w for w in idents if not re.fullmatch(r"[eE]", w)}       # exponent of a float literal
            if any((w not in _MATH_OK) and not (w in names and names[w].kind == "double") for w in idents):
                return [full] * n                       # x used whole, dx on the right, matrices, temporaries, ...
            mask[i] = deps
            continue
        m = re.match(r"^dx\s*=(?!=)\s*([A-Za-z_][A-Za-z0-9_]*)\s*\*\s*x$", st)
        if m and m.group(1) in names and names[m.group(1)].shape == (n, n) and all(v is None for v in mask) and len(stmts) == 1:
            c = names[m.group(1)]
            return [sum(1 << k for k in range(n) if c.slots[i, k] >= 0 or c.values[i, k] != 0.0) for i in range(n)]
        return [full] * n
    return [0 if v is None else v for v in mask]        # a component never assigned stays 0 (dx is zero-initialised)


def modelTemplate(srUpdate: bool = True) -> str:
    """modelTemplate (R/modelTemplate.R:5-1011): the source template of a model translation unit.  The reference's template is
    Rcpp/Armadillo/odeint C++ with the person x time loop written out; here it is a CUDA TU whose filter lives in
    psy_filter.cuh and whose model struct receives the user's equations.  LATENTEQUATIONPLACEHOLDER and
    MEASUREMENTEQUATIONPLACEHOLDER are replaced exactly as newPsydiff does (R/prepareModel.R:509-512); the remaining
    upper-case placeholders carry the compile-time facts the reference reads from the model list at run time."""
    return (
        "// generated by psydiff_b200.codegen — model translation unit for sm_100a\n"
        '#include "psy_filter.cuh"\n\n'
        "struct PsyUserModel {\n"
        "  static constexpr int N = NLATENTPLACEHOLDER, M = NMANIFESTPLACEHOLDER, NSLOT = NSLOTPLACEHOLDER, NFREE = NFREEPLACEHOLDER;\n"
        "  static constexpr int L_STRUCT = LSTRUCTPLACEHOLDER;\n"
        "  static constexpr bool R_DIAG = RDIAGPLACEHOLDER;\n"
        f"  static constexpr bool SR = {'true' if srUpdate else 'false'};\n"
        "  static constexpr int STEPPER = STEPPERPLACEHOLDER;\n"
        "  // bit r of dep(i): drift component i may read x_r (Jacobian sparsity read off the equation text)\n"
        "  static __host__ __device__ constexpr unsigned dep(int i) { return DEPPLACEHOLDER; }\n\n"
        "LATENTEQUATIONPLACEHOLDER\n"
        "MEASUREMENTEQUATIONPLACEHOLDER\n"
        "INITIALPLACEHOLDER"
        "};\n\nPSY_INSTANTIATE(PsyUserModel)\n")


def prepareEquations(equations: str, spec: ModelSpec) -> str:
    """prepareEquations (R/prepareModel.R:737-770): declarations of every parameter container the text refers to, followed
    by the user's statements.  The reference emits `arma::mat X = modelPars["X"];` lookups; here the containers are
    fixed-size values whose free entries read the instance's slot array."""
    out = [_bind(spec.container(name)) for name in referenced(equations, spec)]
    out.append("    " + equations.strip().replace("\n", "\n    ") + "\n")
    return "".join(out)


def emit_cuda(spec: ModelSpec) -> str:
    n, m, F = spec.nlatent, spec.nmanifest, spec.n_slots
    A0, Ll, Rl, m0 = (spec.container(k) for k in ("A0LogChol", "LLogChol", "RLogChol", "m0"))
    if A0.shape != (n, n) or Ll.shape != (n, n):
        raise ValueError("A0 and L must be nlatent x nlatent")
    if Rl.shape != (m, m):
        raise ValueError("Rchol must be nmanifest x nmanifest")
    if m0.shape[0] * m0.shape[1] != n:
        raise ValueError("m0 must have nlatent elements")
    _, a0_lower = _structure(A0)
    if not a0_lower:
        raise ValueError("A0 must be lower triangular (it is used as the Cholesky root of the initial covariance); "
                         "entries above the diagonal are unsupported on device")
    l_diag, _ = _structure(Ll)
    r_diag, _ = _structure(Rl)
    dep = drift_dependency(spec)
    chain = " : ".join(f"i == {i} ? {hex(dep[i])}u" for i in range(n - 1)) + (" : " if n > 1 else "") + f"{hex(dep[n - 1])}u"
    # latentEquation (R/prepareModel.R:493-505) / measurementEquation (R/prepareModel.R:479-491)
    latent = ("  static __device__ __forceinline__ void latentEquation(const psy::Vec<N>& x, psy::Vec<N>& dx,\n"
              "      const double (&psy_p_)[NFREE], const int person, const double t) {\n"
              + prepareEquations(spec.latentEquations, spec) + "  }\n")
    manifest = ("  static __device__ __forceinline__ void measurementEquation(const psy::Vec<N>& x, psy::Vec<M>& y,\n"
                "      const double (&psy_p_)[NFREE], const int person, const double t) {\n"
                + prepareEquations(spec.manifestEquations, spec) + "  }\n")
    # initial state: m0, A0LogChol, LLogChol as stored (log-Cholesky; the kernel applies logChol2Chol_C); Rchol = logChol2Chol(RLogChol)
    ini = ["  static __device__ __forceinline__ void initial(double (&m0)[N], double (&A0lc)[N * N], double (&Llc)[N * N],\n"
           "      const double (&psy_p_)[NFREE]) {\n"]
    m0v, m0s = m0.values.reshape(-1, order="F"), m0.slots.reshape(-1, order="F")
    for i in range(n):
        e = f"psy_p_[{int(m0s[i])}]" if m0s[i] >= 0 else lit(m0v[i])
        ini.append(f"    m0[{i}] = {e};\n")
    for j in range(n):
        for i in range(n):
            ini.append(f"    A0lc[{i + j * n}] = {_elem(A0, i, j)};\n")
    for j in range(n):
        for i in range(n):
            ini.append(f"    Llc[{i + j * n}] = {_elem(Ll, i, j)};\n")
    ini.append("  }\n\n")
    ini.append("  static __device__ __forceinline__ void rchol(double (&Rc)[M * M], const double (&psy_p_)[NFREE]) {\n")
    for j in range(m):
        for i in range(m):
            e = _elem(Rl, i, j)
            ini.append(f"    Rc[{i + j * m}] = {'exp(' + e + ')' if i == j else e};\n")
    ini.append("  }\n")
    code = modelTemplate(spec.srUpdate)
    for key, val in (("NLATENTPLACEHOLDER", str(n)), ("NMANIFESTPLACEHOLDER", str(m)), ("NSLOTPLACEHOLDER", str(F)),
                     ("NFREEPLACEHOLDER", str(max(F, 1))), ("LSTRUCTPLACEHOLDER", "psy::L_DIAG" if l_diag else "psy::L_FULL"),
                     ("RDIAGPLACEHOLDER", "true" if r_diag else "false"), ("STEPPERPLACEHOLDER", str(int(spec.stepper))),
                     ("DEPPLACEHOLDER", chain), ("MEASUREMENTEQUATIONPLACEHOLDER", manifest),
                     ("LATENTEQUATIONPLACEHOLDER", latent), ("INITIALPLACEHOLDER", "".join(ini))):
        code = code.replace(key, val)      # stringr::str_replace_all, R/prepareModel.R:511-512
    return code
