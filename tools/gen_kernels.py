#!/usr/bin/env python3
"""Generate the smoothing-kernel code for the oracle and for the CUDA product.

The reference builds `tit/sph/kernel.inl.hpp` at configure time from the six
piecewise-polynomial kernel definitions with SymEngine
(/root/reference/source/tit/sph/kernel.gen.cpp:520-554, generator 313-478). That
header is not in the reference tree, so it is re-derived here with sympy from
the same definitions.

Two *independent evaluation strategies* are emitted from the one symbolic
source, so that the oracle and the product do not share arithmetic:

  --oracle  PATH   expanded multivariate Horner + CSE straight-line code, the
                   same strategy as the reference generator
                   (kernel.gen.cpp:453-473). Plain C++ (`inline double`).
  --product PATH   runtime-free unrolled J/K recurrences with numeric
                   coefficients (kernel.gen.cpp:313-334 restated as code
                   instead of being expanded symbolically). `__host__
                   __device__` CUDA.

Nothing under /root/reference is read at run time.
"""
from __future__ import annotations

import argparse
from fractions import Fraction

import sympy as sp

q, z, eta, rho, delta, beta, A, L = sp.symbols("q z eta rho delta beta A L")

R = sp.Rational


def kernel_defs():
    """(name, [(cutoff, w(q)), ...]) — kernel.gen.cpp:520-554."""
    return [
        ("CubicSpline", [(R(2), R(1, 4) * (2 - q) ** 3), (R(1), -((1 - q) ** 3))]),
        (
            "QuarticSpline",
            [
                (R(5, 2), (R(5, 2) - q) ** 4),
                (R(3, 2), -5 * (R(3, 2) - q) ** 4),
                (R(1, 2), 10 * (R(1, 2) - q) ** 4),
            ],
        ),
        (
            "QuinticSpline",
            [(R(3), (3 - q) ** 5), (R(2), -6 * (2 - q) ** 5), (R(1), 15 * (1 - q) ** 5)],
        ),
        ("QuarticWendland", [(R(2), (1 + 2 * q) * (1 - q / 2) ** 4)]),
        (
            "SixthOrderWendland",
            [(R(2), (1 + 3 * q + R(35, 12) * q**2) * (1 - q / 2) ** 6)],
        ),
        (
            "EighthOrderWendland",
            [(R(2), (1 + 4 * q + R(25, 4) * q**2 + 4 * q**3) * (1 - q / 2) ** 8)],
        ),
    ]


def coeffs(expr, var):
    """{power: coeff} of the expanded polynomial, ascending."""
    p = sp.Poly(sp.expand(expr), var)
    out = {}
    for (k,), c in p.terms():
        if c != 0:
            out[int(k)] = c
    return dict(sorted(out.items()))


# --- symbolic pieces (kernel.gen.cpp:340-451) --------------------------------


def deriv_expr(w, cutoff):
    """w'(q) in Horner form around the cutoff (kernel.gen.cpp:358-365)."""
    d = sp.expand(sp.diff(w, q))
    t = sp.Symbol("t")
    shifted = sp.horner(sp.expand(d.subs(q, t + cutoff)), t)
    return shifted.subs(t, q - cutoff)


def tail_moment_expr(w, cutoff, dim):
    """Integral of xi^(dim-1) w(xi) from q to cutoff (kernel.gen.cpp:369-375)."""
    r = sum(c * q**p / (dim + p) for p, c in coeffs(w, q).items())
    r = sp.horner(sp.expand(r), q)
    return cutoff**dim * r.subs(q, cutoff) - q**dim * r


def tail_moment_poly(w, cutoff, dim):
    return coeffs(sp.expand(tail_moment_expr(w, cutoff, dim)), q)


def j_expr(p, b2):
    """J_p, primitive of rho^p along a line (kernel.gen.cpp:313-327)."""
    if p == 0:
        return z
    if p == 1:
        return (z * rho + b2 * L) / 2
    return (z * rho**p + p * b2 * j_expr(p - 2, b2)) / (p + 1)


def k_expr(p):
    """K_p, edge primitive (kernel.gen.cpp:329-334)."""
    if p == 0:
        return A
    if p == 1:
        return delta * L
    return delta * j_expr(p - 2, beta**2) + eta**2 * k_expr(p - 2)


def flux_moment(w, upper):
    """kernel.gen.cpp:443-451."""
    return sum(c * (upper ** (p + 2) - eta ** (p + 2)) / (p + 2) for p, c in coeffs(w, q).items())


def seg_flux_sym(w):
    return sum(c * j_expr(p, eta**2) for p, c in coeffs(w, q).items())


def seg_antigrad_sym(w, cutoff):
    out = 0
    for p, c in tail_moment_poly(w, cutoff, 2).items():
        if p == 0:
            out += c * A
        elif p == 1:
            out += c * eta * L
        else:
            out += c * eta * j_expr(p - 2, eta**2)
    return out


def tri_flux_line_sym(w):
    return sum(c * k_expr(p) for p, c in coeffs(flux_moment(w, rho), rho).items())


def tri_antigrad_line_sym(w, cutoff):
    out = 0
    for p, c in tail_moment_poly(w, cutoff, 3).items():
        if p == 0:
            out += c * k_expr(0)
        else:
            out += c * (eta * k_expr(p - 1) - eta**p * k_expr(0)) / (p - 1)
    return out


def tri_flux_sector_sym(w, cutoff):
    return flux_moment(w, cutoff)


def tri_antigrad_sector_sym(w, cutoff):
    out = 0
    for p, c in tail_moment_poly(w, cutoff, 3).items():
        if p == 0:
            out += c * (1 - eta / cutoff)
        else:
            out += c * (eta * cutoff ** (p - 1) - eta**p) / (p - 1)
    return out


def weight(pieces, dim):
    m0 = sum(tail_moment_expr(w, c, dim).subs(q, 0) for c, w in pieces)
    area = {1: 2, 2: 2 * sp.pi, 3: 4 * sp.pi}[dim]
    return 1 / (m0 * area)


# --- printing ----------------------------------------------------------------


def lit(c) -> str:
    """Exact rational -> C double expression."""
    c = sp.sympify(c)
    if c.is_Integer:
        return f"{int(c)}.0"
    if c.is_Rational:
        return f"({int(c.p)}.0 / {int(c.q)}.0)"
    return repr(float(c))


class CPrinter:
    """Minimal C printer: integer powers become `ipow<N>(x)`."""

    def doprint(self, e) -> str:
        e = sp.sympify(e)
        if e.is_Integer or e.is_Rational:
            return lit(e)
        if e.is_Symbol:
            return e.name
        if e.is_Add:
            return "(" + " + ".join(self.doprint(a) for a in e.as_ordered_terms()) + ")"
        if e.is_Mul:
            return "(" + " * ".join(self.doprint(a) for a in e.as_ordered_factors()) + ")"
        if e.is_Pow:
            b, ex = e.as_base_exp()
            if ex.is_Integer and int(ex) >= 1:
                return f"ipow<{int(ex)}>({self.doprint(b)})"
            if ex.is_Integer and int(ex) <= -1:
                return f"(1.0 / ipow<{-int(ex)}>({self.doprint(b)}))"
        raise ValueError(f"cannot print {e!r}")


PR = CPrinter()


def power_reduce(e, sym, squared):
    """sym^n -> squared^(n//2) * sym^(n%2) (kernel.gen.cpp:263-290)."""
    e = sp.expand(e)
    return e.replace(
        lambda x: x.is_Pow and x.base == sym and x.exp.is_Integer and int(x.exp) >= 2,
        lambda x: squared ** (int(x.exp) // 2) * sym ** (int(x.exp) % 2),
    )


def multi_horner(e, vars_):
    """Multivariate Horner (kernel.gen.cpp:240-252)."""
    if not vars_:
        return sp.sympify(e)
    var, rest = vars_[0], vars_[1:]
    cs = coeffs(e, var) if sp.sympify(e).has(var) else {0: e}
    deg = max(cs)
    result = multi_horner(cs.get(deg, 0), rest)
    for p in range(deg - 1, -1, -1):
        result = multi_horner(cs.get(p, 0), rest) + var * result
    return result


def emit_expanded(name, params, expr, opt_vars, reduces) -> str:
    """Expanded Horner + CSE function body (oracle form)."""
    for sym, sq in reduces:
        expr = power_reduce(expr, sym, sq)
    expr = sp.expand(expr)
    # Linear in the transcendental symbols: Horner each coefficient separately.
    lin = [s for s in (rho, A, L) if s in params]
    poly = sp.Poly(expr, *lin) if lin else None
    total = 0
    if poly is not None:
        for mon, c in poly.terms():
            term = multi_horner(c, opt_vars)
            for s, k in zip(lin, mon):
                term = term * s**k
            total += term
    else:
        total = multi_horner(expr, opt_vars)
    reps, (res,) = sp.cse([total], symbols=sp.numbered_symbols("x_"), order="none")
    used = total.free_symbols
    sig = ", ".join(f"double {p.name}" if p in used else f"double /*{p.name}*/" for p in params)
    lines = [f"inline double {name}({sig}) {{"]
    for s, v in reps:
        lines.append(f"  const double {s.name} = {PR.doprint(v)};")
    lines.append(f"  return {PR.doprint(res)};")
    lines.append("}")
    return "\n".join(lines)


# --- recurrence (product) form -------------------------------------------------


def horner_eta(poly: dict[int, sp.Expr]) -> str:
    """Horner in eta of {power: rational}."""
    if not poly:
        return "0.0"
    deg = max(poly)
    s = lit(poly.get(deg, 0))
    for p in range(deg - 1, -1, -1):
        c = poly.get(p, 0)
        s = f"fma(eta, {s}, {lit(c)})" if c != 0 else f"(eta * {s})"
    return s


def emit_recurrence_seg(name, terms: dict[int, sp.Expr], kind: str) -> str:
    """2D segment primitive as unrolled J recurrences.

    kind == "flux":      sum_p c_p J_p
    kind == "antigrad":  c_0 A + c_1 eta L + sum_{p>=2} c_p eta J_{p-2}
    """
    pmax = max(terms)
    jmax = pmax if kind == "flux" else max(pmax - 2, 0)
    out = [f"TIT_HD static double {name}(double eta, double z, double rho, double A, double L) {{"]
    out += ["  (void)A; (void)L; (void)rho;", "  const double e2 = eta * eta;", "  const double r2 = fma(z, z, e2);"]
    # rho^p for p = 0..jmax: even powers from r2, odd ones times rho.
    out.append("  double rp[%d]; rp[0] = 1.0; rp[1] = rho;" % (max(jmax, 1) + 1))
    for p in range(2, jmax + 1):
        out.append(f"  rp[{p}] = rp[{p - 2}] * r2;")
    out.append("  double J[%d];" % (max(jmax, 1) + 1))
    out.append("  J[0] = z;")
    out.append("  J[1] = 0.5 * fma(z, rho, e2 * L);")
    for p in range(2, jmax + 1):
        out.append(f"  J[{p}] = fma(z, rp[{p}], {p}.0 * e2 * J[{p - 2}]) * (1.0 / {p + 1}.0);")
    acc = []
    for p, c in terms.items():
        if kind == "flux":
            acc.append(f"{lit(c)} * J[{p}]")
        elif p == 0:
            acc.append(f"{lit(c)} * A")
        elif p == 1:
            acc.append(f"{lit(c)} * eta * L")
        else:
            acc.append(f"{lit(c)} * eta * J[{p - 2}]")
    out.append("  return " + "\n       + ".join(acc) + ";")
    out.append("}")
    return "\n".join(out)


def emit_recurrence_line(name, w, cutoff, kind: str) -> str:
    """3D triangle edge primitive as unrolled J/K recurrences."""
    if kind == "flux":
        mom = coeffs(flux_moment(w, rho), rho)  # coeffs may depend on eta
        kmax = max(mom)
    else:
        tm = tail_moment_poly(w, cutoff, 3)
        kmax = max(max(tm) - 1, 1)
    jmax = max(kmax - 2, 1)
    out = [
        f"TIT_HD static double {name}(double eta, double delta, double z, double rho, double A, double L) {{",
        "  const double e2 = eta * eta;",
        "  const double b2 = fma(delta, delta, e2);",
        "  const double r2 = fma(z, z, b2);",
        "  double rp[%d]; rp[0] = 1.0; rp[1] = rho;" % (jmax + 1),
    ]
    for p in range(2, jmax + 1):
        out.append(f"  rp[{p}] = rp[{p - 2}] * r2;")
    out.append("  double J[%d];" % (jmax + 1))
    out.append("  J[0] = z;")
    out.append("  J[1] = 0.5 * fma(z, rho, b2 * L);")
    for p in range(2, jmax + 1):
        out.append(f"  J[{p}] = fma(z, rp[{p}], {p}.0 * b2 * J[{p - 2}]) * (1.0 / {p + 1}.0);")
    out.append("  double K[%d];" % (kmax + 1))
    out.append("  K[0] = A;")
    out.append("  K[1] = delta * L;")
    for p in range(2, kmax + 1):
        out.append(f"  K[{p}] = fma(delta, J[{p - 2}], e2 * K[{p - 2}]);")
    acc = []
    if kind == "flux":
        for p, c in mom.items():
            if p == 0:
                acc.append(f"({horner_eta(coeffs(c, eta))}) * K[0]")
            else:
                acc.append(f"{lit(c)} * K[{p}]")
    else:
        # c_0 K_0 + sum_p c_p (eta K_{p-1} - eta^p K_0) / (p - 1): gather K_0.
        k0 = {0: tm.get(0, 0)}
        for p, c in tm.items():
            if p == 0:
                continue
            acc.append(f"{lit(c / (p - 1))} * eta * K[{p - 1}]")
            k0[p] = k0.get(p, 0) - c / (p - 1)
        acc.append(f"({horner_eta(k0)}) * K[0]")
    out.append("  return " + "\n       + ".join(acc) + ";")
    out.append("}")
    return "\n".join(out)


def emit_recurrence_line_delta(name, w, cutoff, kind: str) -> str:
    """F(z1) - F(z0) of the 3D edge primitive from ONE pass of the J/K recurrences.

    The recurrences are linear in their seeds (z rho^p, L, A), so the difference
    of two evaluations on the same line (same eta, delta) obeys the same
    recurrences with the seeds replaced by their differences; dA and dL arrive
    as single transcendentals (sph_kernel.cuh, line_prim_delta). Product only:
    halves the arithmetic of the wall pass, same value up to rounding.
    """
    if kind == "flux":
        mom = coeffs(flux_moment(w, rho), rho)
        kmax = max(mom)
    else:
        tm = tail_moment_poly(w, cutoff, 3)
        kmax = max(max(tm) - 1, 1)
    jmax = max(kmax - 2, 1)
    out = [
        f"TIT_HD static double {name}(double eta, double delta, double z1, double rho1, double z0, double rho0, double dA, double dL) {{",
        "  const double e2 = eta * eta;",
        "  const double b2 = fma(delta, delta, e2);",
        "  const double s1 = fma(z1, z1, b2), s0 = fma(z0, z0, b2);",
        "  double p1[%d], p0[%d]; p1[0] = 1.0; p0[0] = 1.0; p1[1] = rho1; p0[1] = rho0;" % (jmax + 1, jmax + 1),
    ]
    for p in range(2, jmax + 1):
        out.append(f"  p1[{p}] = p1[{p - 2}] * s1; p0[{p}] = p0[{p - 2}] * s0;")
    out.append("  double J[%d];" % (jmax + 1))
    out.append("  J[0] = z1 - z0;")
    out.append("  J[1] = 0.5 * fma(b2, dL, fma(z1, rho1, -(z0 * rho0)));")
    for p in range(2, jmax + 1):
        out.append(f"  J[{p}] = fma({p}.0 * b2, J[{p - 2}], fma(z1, p1[{p}], -(z0 * p0[{p}]))) * (1.0 / {p + 1}.0);")
    out.append("  double K[%d];" % (kmax + 1))
    out.append("  K[0] = dA;")
    out.append("  K[1] = delta * dL;")
    for p in range(2, kmax + 1):
        out.append(f"  K[{p}] = fma(delta, J[{p - 2}], e2 * K[{p - 2}]);")
    acc = []
    if kind == "flux":
        for p, c in mom.items():
            if p == 0:
                acc.append(f"({horner_eta(coeffs(c, eta))}) * K[0]")
            else:
                acc.append(f"{lit(c)} * K[{p}]")
    else:
        k0 = {0: tm.get(0, 0)}
        for p, c in tm.items():
            if p == 0:
                continue
            acc.append(f"{lit(c / (p - 1))} * eta * K[{p - 1}]")
            k0[p] = k0.get(p, 0) - c / (p - 1)
        acc.append(f"({horner_eta(k0)}) * K[0]")
    out.append("  return " + "\n       + ".join(acc) + ";")
    out.append("}")
    return "\n".join(out)


def emit_sector(name, expr, qual) -> str:
    return f"{qual} double {name}(double eta) {{\n  return {horner_eta(coeffs(expr, eta))};\n}}"


def piecewise_sum(pieces, fn) -> str:
    parts = []
    for c, w in pieces:
        parts.append(f"(q < {lit(c)} ? {PR.doprint(fn(w, c))} : 0.0)")
    return " + ".join(parts)


# --- file assembly -------------------------------------------------------------

HEADER_NOTE = "// Generated by tools/gen_kernels.py — do not edit.\n"


def gen_common_struct_body(name, pieces, qual) -> list[str]:
    """value / deriv / moments / weights — same closed forms for both outputs
    (they are the kernel *definition*; kernel.gen.cpp:725-777)."""
    out = []
    rad = max(c for c, _ in pieces)
    out.append(f"  static constexpr int num_pieces = {len(pieces)};")
    out.append(f"  static constexpr double unit_radius = {lit(rad)};")
    for d in (1, 2, 3):
        out.append(f"  static constexpr double weight{d} = {float(weight(pieces, d))!r}; // {weight(pieces, d)}")
    out.append(f"  {qual} double cutoff(int i) {{ constexpr double c[] = {{{', '.join(lit(c) for c, _ in pieces)}}}; return c[i]; }}")
    out.append(f"  {qual} double unit_value(double q) {{\n    return {piecewise_sum(pieces, lambda w, c: w)};\n  }}")
    out.append(f"  {qual} double unit_deriv(double q) {{\n    return {piecewise_sum(pieces, deriv_expr)};\n  }}")
    for d in (1, 2, 3):
        out.append(
            f"  {qual} double unit_moment{d}(double q) {{\n    return {piecewise_sum(pieces, lambda w, c, d=d: tail_moment_expr(w, c, d))};\n  }}"
        )
    return out


def generate_oracle(path):
    out = [HEADER_NOTE, "#pragma once", "// Oracle form: expanded Horner + CSE (test infrastructure only).", "#include <cmath>", ""]
    out.append("namespace oracle_gen {")
    out.append("using std::fma;")
    out.append("template<int N> inline double ipow(double x) { if constexpr (N == 1) return x; else { const double h = ipow<N / 2>(x); if constexpr (N % 2) return h * h * x; else return h * h; } }")
    for kid, (name, pieces) in enumerate(kernel_defs()):
        out.append(f"\n// ---- {name} ----")
        out.append(f"namespace k{kid} {{")
        for i, (c, w) in enumerate(pieces):
            out.append(emit_expanded(f"seg_flux_{i}", [eta, z, rho, A, L], seg_flux_sym(w), [z, eta], [(rho, z**2 + eta**2)]))
            out.append(emit_expanded(f"seg_antigrad_{i}", [eta, z, rho, A, L], seg_antigrad_sym(w, c), [z, eta], [(rho, z**2 + eta**2)]))
            red = [(rho, z**2 + eta**2 + delta**2), (beta, eta**2 + delta**2)]
            out.append(emit_expanded(f"tri_flux_line_{i}", [eta, delta, z, rho, A, L], tri_flux_line_sym(w), [z, delta, eta], red))
            out.append(emit_expanded(f"tri_antigrad_line_{i}", [eta, delta, z, rho, A, L], tri_antigrad_line_sym(w, c), [z, delta, eta], red))
            out.append(emit_sector(f"tri_flux_sector_{i}", tri_flux_sector_sym(w, c), "inline"))
            out.append(emit_sector(f"tri_antigrad_sector_{i}", tri_antigrad_sector_sym(w, c), "inline"))
        out.append("} // namespace")
        out.append(f"struct K{kid} {{ // {name}")
        out.append(f'  static constexpr const char* name = "{name}";')
        out += gen_common_struct_body(name, pieces, "static inline")
        for fn, params in (
            ("seg_flux", "double eta, double z, double rho, double A, double L"),
            ("seg_antigrad", "double eta, double z, double rho, double A, double L"),
            ("tri_flux_line", "double eta, double delta, double z, double rho, double A, double L"),
            ("tri_antigrad_line", "double eta, double delta, double z, double rho, double A, double L"),
            ("tri_flux_sector", "double eta"),
            ("tri_antigrad_sector", "double eta"),
        ):
            args = ", ".join(p.split()[-1] for p in params.split(","))
            body = " ".join(f"if (i == {i}) return k{kid}::{fn}_{i}({args});" for i in range(len(pieces)))
            out.append(f"  static inline double {fn}(int i, {params}) {{ {body} return 0.0; }}")
        out.append("};")
    out.append("} // namespace oracle_gen")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def generate_product(path):
    out = [HEADER_NOTE, "#pragma once", "// Product form: unrolled J/K recurrences, host+device.", ""]
    out.append("#ifndef TIT_HD\n#ifdef __CUDACC__\n#define TIT_HD __host__ __device__ __forceinline__\n#else\n#define TIT_HD inline\n#endif\n#endif")
    out.append("namespace titgpu_gen {")
    out.append("template<int N> TIT_HD double ipow(double x) { if constexpr (N == 1) return x; else { const double h = ipow<N / 2>(x); if constexpr (N % 2) return h * h * x; else return h * h; } }")
    out.append("template<int KernelId> struct KernelGen;")
    for kid, (name, pieces) in enumerate(kernel_defs()):
        out.append(f"\n// ---- {name} ----")
        out.append(f"template<> struct KernelGen<{kid}> {{")
        out += gen_common_struct_body(name, pieces, "TIT_HD static")
        for i, (c, w) in enumerate(pieces):
            out.append(emit_recurrence_seg(f"seg_flux_{i}", coeffs(w, q), "flux"))
            out.append(emit_recurrence_seg(f"seg_antigrad_{i}", tail_moment_poly(w, c, 2), "antigrad"))
            out.append(emit_recurrence_line(f"tri_flux_line_{i}", w, c, "flux"))
            out.append(emit_recurrence_line(f"tri_antigrad_line_{i}", w, c, "antigrad"))
            out.append(emit_recurrence_line_delta(f"tri_flux_line_delta_{i}", w, c, "flux"))
            out.append(emit_recurrence_line_delta(f"tri_antigrad_line_delta_{i}", w, c, "antigrad"))
            out.append(emit_sector(f"tri_flux_sector_{i}", tri_flux_sector_sym(w, c), "TIT_HD static"))
            out.append(emit_sector(f"tri_antigrad_sector_{i}", tri_antigrad_sector_sym(w, c), "TIT_HD static"))
        for fn, params in (
            ("seg_flux", "double eta, double z, double rho, double A, double L"),
            ("seg_antigrad", "double eta, double z, double rho, double A, double L"),
            ("tri_flux_line", "double eta, double delta, double z, double rho, double A, double L"),
            ("tri_antigrad_line", "double eta, double delta, double z, double rho, double A, double L"),
            ("tri_flux_line_delta", "double eta, double delta, double z1, double rho1, double z0, double rho0, double dA, double dL"),
            ("tri_antigrad_line_delta", "double eta, double delta, double z1, double rho1, double z0, double rho0, double dA, double dL"),
            ("tri_flux_sector", "double eta"),
            ("tri_antigrad_sector", "double eta"),
        ):
            args = ", ".join(p.split()[-1] for p in params.split(","))
            body = " ".join(f"if constexpr (I == {i}) return {fn}_{i}({args});" for i in range(len(pieces)))
            out.append(f"  template<int I> TIT_HD static double {fn}({params}) {{ {body} return 0.0; }}")
        out.append("};")
    out.append("} // namespace titgpu_gen")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--oracle")
    ap.add_argument("--product")
    a = ap.parse_args()
    if a.oracle:
        generate_oracle(a.oracle)
    if a.product:
        generate_product(a.product)


if __name__ == "__main__":
    main()
