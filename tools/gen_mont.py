#!/usr/bin/env python3
"""Generator + CPU simulator for the Montgomery field kernels' inline PTX.

Emits typlonk_b200/csrc/mont_gen.cuh: one `asm` block per operation (mul, sqr-as-mul,
add, sub) for Fr (8 x u32 limbs) and Fq (12 x u32 limbs) of BLS12-381, with the modulus
limbs as PTX immediates.  Each operation is first built as a small IR (list of PTX-like
instructions) which this script *simulates* on random and edge-case operands against
Python big-int arithmetic, so the carry chains are proven correct on the CPU before any
GPU time is spent.

Multiplication is an interleaved (CIOS-style) Montgomery product over two accumulator
arrays ("even"/"odd" columns) so that every 32x32->64 product is a mad.lo.cc/madc.hi.cc
pair on an aligned register pair -- the pattern ptxas fuses into IMAD.WIDE.U32(.X).

Usage: python tools/gen_mont.py [--check-only]
"""
import random
import sys
from pathlib import Path

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
MASK = 0xFFFFFFFF


def limbs(x, n):
    return [(x >> (32 * i)) & MASK for i in range(n)]


class Prog:
    """Tiny PTX subset.  Registers are strings; ints are immediates."""

    def __init__(self):
        self.ins = []
        self.temps = []

    def emit(self, op, dst, *src, drop_carry=False):
        self.ins.append((op, dst, src, drop_carry))

    # ---- simulation ---------------------------------------------------------
    def run(self, env):
        cc = 0
        pred = {}

        def val(s):
            return s if isinstance(s, int) else env[s]

        for op, dst, src, drop in self.ins:
            s = [val(x) for x in src if not (isinstance(x, str) and x.startswith("%p"))]
            if op == "mul.lo.u32":
                env[dst] = (s[0] * s[1]) & MASK
            elif op == "mul.hi.u32":
                env[dst] = (s[0] * s[1]) >> 32
            elif op in ("mad.lo.cc.u32", "madc.lo.cc.u32"):
                cin = cc if op.startswith("madc") else 0
                t = ((s[0] * s[1]) & MASK) + s[2] + cin
                env[dst] = t & MASK
                cc = t >> 32
                if drop:
                    assert cc == 0, "dropped carry must be zero"
            elif op in ("mad.hi.cc.u32", "madc.hi.cc.u32", "madc.hi.u32"):
                cin = cc if op.startswith("madc") else 0
                t = ((s[0] * s[1]) >> 32) + s[2] + cin
                env[dst] = t & MASK
                if ".cc" in op:
                    cc = t >> 32
                    if drop:
                        assert cc == 0, "dropped carry must be zero"
                else:
                    assert t >> 32 == 0, "carry out of non-.cc op must be zero"
            elif op in ("add.cc.u32", "addc.cc.u32", "addc.u32", "add.u32"):
                cin = cc if op.startswith("addc") else 0
                t = s[0] + s[1] + cin
                env[dst] = t & MASK
                if ".cc" in op:
                    cc = t >> 32
                    if drop:
                        assert cc == 0, "dropped carry must be zero"
                elif drop:
                    assert t >> 32 == 0, "carry out must be zero"
            elif op in ("sub.cc.u32", "subc.cc.u32", "subc.u32"):
                bin_ = cc if op.startswith("subc") else 0
                t = s[0] - s[1] - bin_
                env[dst] = t & MASK
                if ".cc" in op:
                    cc = 1 if t < 0 else 0
            elif op == "setp.ne.u32":
                pred[dst] = s[0] != s[1]
            elif op == "selp.u32":
                env[dst] = s[0] if pred[src[2]] else s[1]
            elif op == "and.b32":
                env[dst] = s[0] & s[1]
            elif op == "mov.u32":
                env[dst] = s[0]
            else:
                raise ValueError(op)
        return env

    # ---- PTX text -----------------------------------------------------------
    def text(self, regmap):
        def r(x):
            if isinstance(x, int):
                return "0x%08x" % x
            return regmap.get(x, x)

        lines = []
        for op, dst, src, _ in self.ins:
            lines.append("%s %s, %s;" % (op, r(dst), ", ".join(r(x) for x in src)))
        return lines


def cond_sub_p(pr, t, d, out, mod, n):
    """out = t >= p ? t - p : t   (t < 2p < 2^(32n))."""
    pl = limbs(mod, n)
    for i in range(n):
        pr.emit("sub.cc.u32" if i == 0 else "subc.cc.u32", d[i], t[i], pl[i])
    pr.emit("subc.u32", "brw", 0, 0)
    pr.emit("setp.ne.u32", "%pb", "brw", 0)
    for i in range(n):
        pr.emit("selp.u32", out[i], t[i], d[i], "%pb")


def build_mul(mod, n, reduce_final=True, mod_regs=False, m0_reg=False, special=False):
    """special=True replaces products by modulus limbs equal to 1 / 0xffffffff (Fr's two lowest
    limbs) with additions: 16 of the 128 products of an Fr multiplication."""
    pl = ["p%d" % i for i in range(n)] if mod_regs else limbs(mod, n)
    m0 = (-pow(mod, -1, 1 << 32)) & MASK
    if m0_reg:
        m0 = "m0"  # register operand, see main()
    pr = Prog()
    a = ["a%d" % i for i in range(n)]
    b = ["b%d" % i for i in range(n)]
    X = ["e%d" % i for i in range(n)]
    Y = ["o%d" % i for i in range(n)]

    def cmad(acc, src, off, m):
        """acc[j],acc[j+1] += src[j+off]*m for j = 0,2,..; leaves carry in CC."""
        for j in range(0, n, 2):
            lim = src[j + off]
            if special and lim == 1 and j == 0:
                # m * 1: no multiplier needed
                pr.emit("add.cc.u32", acc[j], acc[j], m)
                pr.emit("addc.cc.u32", acc[j + 1], acc[j + 1], 0)
                continue
            if special and lim == MASK and j == 0:
                # m * (2^32 - 1) = (hm : nm) with nm = -m, hm = m - (m != 0), prepared by the caller
                pr.emit("add.cc.u32", acc[j], acc[j], "nm")
                pr.emit("addc.cc.u32", acc[j + 1], acc[j + 1], "hm")
                continue
            pr.emit("mad.lo.cc.u32" if j == 0 else "madc.lo.cc.u32", acc[j], lim, m, acc[j])
            pr.emit("madc.hi.cc.u32", acc[j + 1], lim, m, acc[j + 1], drop_carry=(off == 1 and j == n - 2))

    def mad_n_redc(even, odd, bi, first):
        if first:
            for j in range(0, n, 2):
                pr.emit("mul.lo.u32", odd[j], a[j + 1], bi)
                pr.emit("mul.hi.u32", odd[j + 1], a[j + 1], bi)
            for j in range(0, n, 2):
                pr.emit("mul.lo.u32", even[j], a[j], bi)
                pr.emit("mul.hi.u32", even[j + 1], a[j], bi)
        else:
            pr.emit("add.cc.u32", even[0], even[0], odd[1])
            for j in range(0, n - 2, 2):
                pr.emit("madc.lo.cc.u32", odd[j], a[j + 1], bi, odd[j + 2])
                pr.emit("madc.hi.cc.u32", odd[j + 1], a[j + 1], bi, odd[j + 3])
            pr.emit("madc.lo.cc.u32", odd[n - 2], a[n - 1], bi, 0)
            pr.emit("madc.hi.u32", odd[n - 1], a[n - 1], bi, 0)
            cmad(even, a, 0, bi)
            pr.emit("addc.u32", odd[n - 1], odd[n - 1], 0, drop_carry=True)
        pr.emit("mul.lo.u32", "mi", even[0], m0)
        if special and pl[1] == MASK:
            pr.emit("sub.cc.u32", "nm", 0, "mi")   # borrow <=> mi != 0
            pr.emit("subc.u32", "hm", "mi", 0)
        cmad(odd, pl, 1, "mi")
        cmad(even, pl, 0, "mi")
        pr.emit("addc.u32", odd[n - 1], odd[n - 1], 0, drop_carry=True)

    for i in range(0, n, 2):
        mad_n_redc(X, Y, b[i], first=(i == 0))
        mad_n_redc(Y, X, b[i + 1], first=False)
    # merge: X[i] += Y[i+1]
    pr.emit("add.cc.u32", X[0], X[0], Y[1])
    for i in range(1, n - 1):
        pr.emit("addc.cc.u32", X[i], X[i], Y[i + 1])
    pr.emit("addc.u32", X[n - 1], X[n - 1], 0, drop_carry=True)
    out = ["r%d" % i for i in range(n)]
    if reduce_final:
        cond_sub_p(pr, X, Y, out, mod, n)  # Y reused as scratch for t - p
    else:
        for i in range(n):
            pr.emit("mov.u32", out[i], X[i])
    return pr


def build_add(mod, n):
    pr = Prog()
    a = ["a%d" % i for i in range(n)]
    b = ["b%d" % i for i in range(n)]
    t = ["e%d" % i for i in range(n)]
    d = ["o%d" % i for i in range(n)]
    for i in range(n):
        op = "add.cc.u32" if i == 0 else ("addc.cc.u32" if i < n - 1 else "addc.u32")
        pr.emit(op, t[i], a[i], b[i], drop_carry=(i == n - 1))
    cond_sub_p(pr, t, d, ["r%d" % i for i in range(n)], mod, n)
    return pr


def build_sub(mod, n):
    pl = limbs(mod, n)
    pr = Prog()
    a = ["a%d" % i for i in range(n)]
    b = ["b%d" % i for i in range(n)]
    t = ["e%d" % i for i in range(n)]
    for i in range(n):
        pr.emit("sub.cc.u32" if i == 0 else "subc.cc.u32", t[i], a[i], b[i])
    pr.emit("subc.u32", "brw", 0, 0)  # 0xffffffff on borrow
    for i in range(n):
        pr.emit("and.b32", "o%d" % i, "brw", pl[i])
    for i in range(n):
        op = "add.cc.u32" if i == 0 else ("addc.cc.u32" if i < n - 1 else "addc.u32")
        pr.emit(op, "r%d" % i, t[i], "o%d" % i)
    return pr


def check(name, mod, n, trials=300, mod_regs=False, m0_reg=False, special=False):
    R = 1 << (32 * n)
    rinv = pow(R, -1, mod)
    mul = build_mul(mod, n, mod_regs=mod_regs, m0_reg=m0_reg, special=special)
    add = build_add(mod, n)
    sub = build_sub(mod, n)
    rnd = random.Random(1234 + n)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, R % mod, (1 << (32 * n - 1)) % mod]
    cases = [(x, y) for x in edge for y in edge]
    cases += [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(trials)]
    for x, y in cases:
        env = {}
        for i, v in enumerate(limbs(x, n)):
            env["a%d" % i] = v
        for i, v in enumerate(limbs(y, n)):
            env["b%d" % i] = v
        for i, v in enumerate(limbs(mod, n)):
            env["p%d" % i] = v
        env["m0"] = (-pow(mod, -1, 1 << 32)) & MASK
        out = mul.run(dict(env))
        got = sum(out["r%d" % i] << (32 * i) for i in range(n))
        assert got == x * y * rinv % mod, (name, "mul", hex(x), hex(y))
        out = add.run(dict(env))
        got = sum(out["r%d" % i] << (32 * i) for i in range(n))
        assert got == (x + y) % mod, (name, "add")
        out = sub.run(dict(env))
        got = sum(out["r%d" % i] << (32 * i) for i in range(n))
        assert got == (x - y) % mod, (name, "sub")
    nmad = sum(1 for i in mul.ins if i[0].startswith(("mad", "mul")))
    print("%s: %d cases ok; mul = %d PTX instrs (%d mul/mad)" % (name, len(cases), len(mul.ins), nmad))


def emit_fn(fname, pr, n, mod_sym=None, m0_sym=None):
    regmap = {}
    if m0_sym:
        regmap["m0"] = "%%%d" % (3 * n)
    if mod_sym:
        for i in range(n):
            regmap["p%d" % i] = "%%%d" % (3 * n + i)
    for i in range(n):
        regmap["r%d" % i] = "%%%d" % i
        regmap["a%d" % i] = "%%%d" % (n + i)
        regmap["b%d" % i] = "%%%d" % (2 * n + i)
    body = pr.text(regmap)
    lines = []
    lines.append("__device__ __forceinline__ void %s(uint32_t* __restrict__ r, const uint32_t* a, const uint32_t* b) {" % fname)
    lines.append("  asm(\"{\\n\\t\"")
    lines.append("      \".reg .u32 e<%d>, o<%d>, mi, nm, hm, brw;\\n\\t\"" % (n, n))
    lines.append("      \".reg .pred %%pb;\\n\\t\"")
    for ln in body:
        lines.append("      \"%s\\n\\t\"" % ln.replace("%pb", "%%pb"))
    lines.append("      \"}\"")
    outs = ", ".join("\"=r\"(r[%d])" % i for i in range(n))
    ins = ", ".join("\"r\"(a[%d])" % i for i in range(n)) + ", " + ", ".join("\"r\"(b[%d])" % i for i in range(n))
    if mod_sym:
        ins += ", " + ", ".join("\"r\"(%s[%d])" % (mod_sym, i) for i in range(n))
    if m0_sym:
        ins += ", \"r\"(%s)" % m0_sym
    lines.append("      : %s" % outs)
    lines.append("      : %s);" % ins)
    lines.append("}")
    return "\n".join(lines)


def main():
    check("Fr", R_MOD, 8)
    check("Fr(mod in regs)", R_MOD, 8, mod_regs=True)
    check("Fr(m0 in a register)", R_MOD, 8, m0_reg=True)
    check("Fr(m0 in a register, low limbs by addition)", R_MOD, 8, m0_reg=True, special=True, trials=3000)
    check("Fq", Q_MOD, 12)
    if "--check-only" in sys.argv:
        return
    out = []
    out.append("// GENERATED by tools/gen_mont.py -- do not edit.  Inline-PTX Montgomery kernels for\n"
               "// BLS12-381 Fr (8 x u32) and Fq (12 x u32); carry chains validated by the generator's\n"
               "// CPU simulator against big-integer arithmetic.\n#pragma once\n#include <stdint.h>\n")
    def carr(x, n):
        return "{" + ", ".join("0x%08xu" % v for v in limbs(x, n)) + "}"
    for nm, mod, n in (("FR", R_MOD, 8), ("FQ", Q_MOD, 12)):
        R = (1 << (32 * n)) % mod
        out.append("#define TP_%s_MOD %s" % (nm, carr(mod, n)))
        out.append("#define TP_%s_ONE %s   // R mod p" % (nm, carr(R, n)))
        out.append("#define TP_%s_R2 %s   // R^2 mod p" % (nm, carr(R * R % mod, n)))
        out.append("#define TP_%s_MODM2 %s   // p - 2 (Fermat inversion exponent)" % (nm, carr(mod - 2, n)))
    # Fr has -r^-1 mod 2^32 = 0xffffffff.  When ptxas sees that immediate it rewrites mi = -t0 and
    # then splits every reduction product mi * r_j into IMAD.X + IMAD.HI.U32.X (6.6 clk on the
    # FMA-heavy pipe instead of 4 for the fused IMAD.WIDE.U32.X) -- about half of all products of
    # the multiplication.  Feeding the constant through a __constant__ word keeps it opaque: 8 extra
    # 32-bit IMADs, all 128 products fused.  Fq's 0xfffcfffd does not trigger the rewrite.
    out.append("static __constant__ uint32_t TP_FR_M0 = 0xffffffffu;   // -r^-1 mod 2^32, deliberately not an immediate")
    for name, mod, n in (("fr", R_MOD, 8), ("fq", Q_MOD, 12)):
        if name == "fr":
            out.append(emit_fn("fr_mul_ptx", build_mul(mod, n, m0_reg=True, special=True), n, m0_sym="TP_FR_M0"))
        else:
            out.append(emit_fn("%s_mul_ptx" % name, build_mul(mod, n), n))
        out.append(emit_fn("%s_add_ptx" % name, build_add(mod, n), n))
        out.append(emit_fn("%s_sub_ptx" % name, build_sub(mod, n), n))
    dst = Path(__file__).resolve().parent.parent / "typlonk_b200" / "csrc" / "mont_gen.cuh"
    dst.parent.mkdir(parents=True, exist_ok=True)
    dst.write_text("\n\n".join(out) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
