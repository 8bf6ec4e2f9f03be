#!/usr/bin/env python3
"""Generate frequensea_b200/csrc/fsea_codelets.cuh: straight-line forward-DFT
register codelets of size R in {2,4,8,16,32,64}, type-generic over the vector
type V (float, double, or the packed 2 x fp32 pair f2 -> FADD2/FMUL2/FFMA2).

Structure: recursive radix-2 decimation in time.  Every butterfly
    X0 = E + w*O,  X1 = E - w*O,   w = exp(-2 pi i k / L) = c - i s
is emitted in "tangent" form so that it costs 6 fused ops instead of 8:
    |c| >= |s|:  u = O * (1 - i t), t = s/c ;  X0 = E + c u ; X1 = E - c u
    |c| <  |s|:  u = O * (t - i),   t = c/s ;  X0 = E + s u ; X1 = E - s u
with the trivial twiddles (1, -i) costing 4 adds.  Twiddle constants are
evaluated here in extended precision and emitted as literals (the compiler
folds them into FFMA/FFMA2 immediates).

    python tools/gen_codelets.py > frequensea_b200/csrc/fsea_codelets.cuh
"""
import sys
from decimal import Decimal, getcontext
from fractions import Fraction
import math

getcontext().prec = 50


def _pi():
    # Machin-like via decimal for > double precision
    getcontext().prec += 5
    def arctan_inv(x):
        x = Decimal(x); s = Decimal(0); t = 1 / x; x2 = x * x; n = 1; sign = 1
        while t / n > Decimal(10) ** (-getcontext().prec):
            s += sign * t / n; t /= x2; n += 2; sign = -sign
        return s
    pi = 4 * (4 * arctan_inv(5) - arctan_inv(239))
    getcontext().prec -= 5
    return +pi

PI = _pi()


def _cos_sin(frac: Fraction):
    """cos/sin of 2*pi*frac in Decimal."""
    x = 2 * PI * Decimal(frac.numerator) / Decimal(frac.denominator)
    # Taylor (|x| <= 2*pi, 50 digits: fine)
    def series(first, start):
        term = first; s = term; n = start
        while abs(term) > Decimal(10) ** (-45):
            term = -term * x * x / (n * (n + 1)); s += term; n += 2
        return s
    return series(Decimal(1), 1), series(x, 2)


class Gen:
    def __init__(self):
        self.lines = []
        self.n = 0
        self.ops = 0

    def tmp(self):
        self.n += 1
        return f"t{self.n}"

    def emit(self, expr):
        v = self.tmp()
        self.lines.append(f"    const V {v} = {expr};")
        self.ops += 1
        return v

    @staticmethod
    def lit(d: Decimal):
        return f"S({d:.25E})"

    def add(self, a, b): return self.emit(f"vadd({a}, {b})")
    def sub(self, a, b): return self.emit(f"vsub({a}, {b})")
    def fma(self, a, c, b): return self.emit(f"vfma({a}, {self.lit(c)}, {b})")     # a*c + b
    def fnma(self, a, c, b): return self.emit(f"vfnma({a}, {self.lit(c)}, {b})")   # b - a*c
    def fms(self, a, c, b): return self.emit(f"vfms({a}, {self.lit(c)}, {b})")     # a*c - b

    # ---- one butterfly, split in two phases so that a stage can be emitted
    # ---- "all first halves, then all second halves" (independent neighbours)
    def bfly_pre(self, O, k, L):
        """Phase 1: the twiddled odd input, up to the final scale. Returns a closure state."""
        orr, oi = O
        fr = Fraction(k, L)
        if k == 0:
            return ("one", orr, oi, None)
        if fr == Fraction(1, 4):                  # w = -i : w*O = (oi, -or)
            return ("mi", orr, oi, None)
        c, s = _cos_sin(fr)                       # w = c - i s ; w*O = (c or + s oi, c oi - s or)
        if abs(c) >= abs(s) or abs(abs(c) - abs(s)) < Decimal(10) ** (-40):
            t = s / c
            if abs(abs(t) - 1) < Decimal(10) ** (-40):
                if t > 0:
                    ur, ui = self.add(orr, oi), self.sub(oi, orr)
                else:
                    ur, ui = self.sub(orr, oi), self.add(oi, orr)
            else:
                ur = self.fma(oi, t, orr)         # or + t*oi
                ui = self.fnma(orr, t, oi)        # oi - t*or
            m = c
        else:
            t = c / s
            ur = self.fma(orr, t, oi)             # t*or + oi
            ui = self.fms(oi, t, orr)             # t*oi - or
            m = s
        return ("gen", ur, ui, m)

    def bfly_post(self, E, st):
        er, ei = E
        kind, ur, ui, m = st
        if kind == "one":
            return (self.add(er, ur), self.add(ei, ui)), (self.sub(er, ur), self.sub(ei, ui))
        if kind == "mi":                          # (ur, ui) = (or, oi); w*O = (oi, -or)
            return (self.add(er, ui), self.sub(ei, ur)), (self.sub(er, ui), self.add(ei, ur))
        x0 = (self.fma(ur, m, er), self.fma(ui, m, ei))
        x1 = (self.fnma(ur, m, er), self.fnma(ui, m, ei))
        return x0, x1

    def addc(self, a, c): return self.emit(f"vaddc({a}, {self.lit(c)})")             # a + c

    def fft(self, xs, chunk=None, raw_bias=None):
        import os
        chunk = chunk or int(os.environ.get("FSEA_CODELET_CHUNK", "4"))
        """Iterative radix-2 DIT, emitted stage by stage (breadth first): within a
        stage the butterflies are independent, and they are written out `chunk` at
        a time -- first every butterfly's twiddle half, then every butterfly's
        add/sub half -- so that neighbouring instructions never depend on each
        other (the packed FFMA2 pipe has few warps to hide its latency behind)."""
        R = len(xs)
        lg = R.bit_length() - 1
        v = [xs[int(format(j, f"0{lg}b")[::-1], 2)] if lg else xs[j] for j in range(R)]
        half = 1
        while half < R:
            L = 2 * half
            pairs = [(base + k, base + k + half, k) for base in range(0, R, L) for k in range(half)]
            # trivial-twiddle butterflies first, then the twiddled ones, each in chunks
            pairs.sort(key=lambda p: (0 if (p[2] == 0 or Fraction(p[2], L) == Fraction(1, 4)) else 1))
            if half == 1 and raw_bias is not None:
                # inputs are raw "magic" floats f = bias + c: the difference needs no
                # unpacking at all (the bias cancels), the sum loses 2*bias in one add
                for (a, b, k) in pairs:
                    (er, ei), (orr, oi) = v[a], v[b]
                    x0 = (self.addc(self.add(er, orr), -2 * raw_bias), self.addc(self.add(ei, oi), -2 * raw_bias))
                    x1 = (self.sub(er, orr), self.sub(ei, oi))
                    v[a], v[b] = x0, x1
                half = L
                continue
            for c0 in range(0, len(pairs), chunk):
                grp = pairs[c0:c0 + chunk]
                pre = [self.bfly_pre(v[b], k, L) for (_, b, k) in grp]
                for (a, b, k), st in zip(grp, pre):
                    v[a], v[b] = self.bfly_post(v[a], st)
            half = L
        return v


RAW_BIAS = Decimal(32896)      # 0x47000000 | (byte << 8) = 32768 + byte ; centred = f - 32896


def codelet(R, raw=False):
    g = Gen()
    xs = [(f"re[{j}]", f"im[{j}]") for j in range(R)]
    out = g.fft(xs, raw_bias=RAW_BIAS if raw else None)
    body = "\n".join(g.lines)
    stores = "\n".join(f"    re[{k}] = {o[0]}; im[{k}] = {o[1]};" for k, o in enumerate(out))
    name = f"fsea_dft{R}_raw" if raw else f"fsea_dft{R}"
    note = (" -- inputs are raw magic floats 32768 + byte (8-bit IQ placed in mantissa bits 8..15);\n"
            "// the byte-unpack subtraction is folded into the first butterfly stage") if raw else ""
    return (f"// radix-{R}{' raw' if raw else ''}: {g.ops} fused ops ({g.ops / R:.2f} per point){note}\n"
            f"template <typename V>\nFSEA_HD void {name}(V* __restrict__ re, V* __restrict__ im) {{\n"
            f"    typedef typename fsea_vt<V>::scalar S;\n{body}\n{stores}\n}}\n"), g.ops


def cross(R):
    """X[m] = A_m + W_{2R}^m B_m, X[m+R] = A_m - W_{2R}^m B_m on the two lanes of R packed registers
    (A in .x, B in .y -> X[m] in .x, X[m+R] in .y): the last radix-2 stage of a 2R-point
    decimation in time whose halves were transformed together by one packed radix-R codelet."""
    L = [f"template <> FSEA_D void fsea_cross<{R}>(f2* __restrict__ re, f2* __restrict__ im) {{"]
    for m in range(R):
        L.append(f"    {{ const float ar = re[{m}].v.x, br = re[{m}].v.y, ai = im[{m}].v.x, bi = im[{m}].v.y;")
        fr = Fraction(m, 2 * R)
        if m == 0:
            L.append("      re[0].v.x = ar + br; re[0].v.y = ar - br; im[0].v.x = ai + bi; im[0].v.y = ai - bi; }")
            continue
        if fr == Fraction(1, 4):                      # W = -i : W*B = (bi, -br)
            L.append(f"      re[{m}].v.x = ar + bi; re[{m}].v.y = ar - bi; im[{m}].v.x = ai - br; im[{m}].v.y = ai + br; }}")
            continue
        c, s_ = _cos_sin(fr)                          # W = c - i s ; W*B = (c br + s bi, c bi - s br)
        if abs(c) >= abs(s_):
            t, mm = s_ / c, c
            L.append(f"      const float ur = fmaf({float(t)!r}f, bi, br), ui = fmaf({float(-t)!r}f, br, bi);")
        else:
            t, mm = c / s_, s_
            L.append(f"      const float ur = fmaf({float(t)!r}f, br, bi), ui = fmaf({float(t)!r}f, bi, -br);")
        L.append(f"      re[{m}].v.x = fmaf({float(mm)!r}f, ur, ar); re[{m}].v.y = fmaf({float(-mm)!r}f, ur, ar);"
                 f" im[{m}].v.x = fmaf({float(mm)!r}f, ui, ai); im[{m}].v.y = fmaf({float(-mm)!r}f, ui, ai); }}")
    L.append("}")
    return "\n".join(L) + "\n"


def crosses():
    out = ["// cross-lane radix-2 stage of a 2R-point DIT: in (A_m, B_m), out (X[m], X[m+R]); scalar fp32",
           "#if defined(__CUDACC__) || defined(FSEA_EMU)",
           "template <int R> FSEA_D void fsea_cross(f2* __restrict__ re, f2* __restrict__ im);"]
    for R in (2, 4, 8, 16, 32):
        out.append(cross(R))
    out.append("#endif")
    return "\n".join(out) + "\n"


def main():
    out = ["// GENERATED by tools/gen_codelets.py -- do not edit.\n"
           "// Forward DFT register codelets, X[k] = sum_j x[j] exp(-2 pi i j k / R),\n"
           "// natural order in, natural order out, in place on re[]/im[].\n"
           "#pragma once\n#include \"fsea_vec.cuh\"\n"]
    out.append("template <typename V>\nFSEA_HD void fsea_dft1(V*, V*) {}\n")
    for R in (2, 4, 8, 16, 32, 64):
        src, ops = codelet(R)
        out.append(src)
    for R in (8, 16, 32, 64):
        src, ops = codelet(R, raw=True)
        out.append(src)
    out.append(crosses())
    out.append("template <int R, typename V> struct fsea_dft;\n" + "".join(
        f"template <typename V> struct fsea_dft<{R}, V> {{ static FSEA_HD void run(V* re, V* im) {{ fsea_dft{R}<V>(re, im); }} }};\n"
        for R in (1, 2, 4, 8, 16, 32, 64)))
    out.append("template <int R, typename V> struct fsea_dft_raw;\n" + "".join(
        f"template <typename V> struct fsea_dft_raw<{R}, V> {{ static FSEA_HD void run(V* re, V* im) {{ fsea_dft{R}_raw<V>(re, im); }} }};\n"
        for R in (8, 16, 32, 64)))
    sys.stdout.write("\n".join(out))


if __name__ == "__main__":
    main()
