#!/usr/bin/env python
"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE TEXT.

The reference is Rust and cannot be compiled in this image (no cargo / rustc). Its arithmetic kernels, however, are
straight-line float64 code in a tiny syntax subset (`let`, `+=`, `if`, `.sqrt()`, `.powi(n)`, struct literals). This
script reads those function bodies from /root/reference at GENERATION time, translates them statement by statement
to Python with a mechanical, regex-level translator (no hand transcription of any formula or coefficient), executes
them on seeded inputs and stores inputs + outputs in tests/golden/reference_exec.npz. Python floats are IEEE-754
binary64 with the same correctly rounded + - * / sqrt as Rust's f64 and no fused contraction, so the stored numbers
are what the Rust code computes, bit for bit (`f64::powi` = compiler-rt `__powidf2`, restated below).

Covered (file:line of the reference):
  kernel.rs:41-82    kernel_potential_per_unit_mass / kernel_accel_factor  (wrappers re-stated here: 2 x 10 lines)
  kernel.rs:84-128   w2, w2_prime                                           (translated)
  multipole.rs:82-170    MultipoleMoment::from_points_const  (P2M)          (translated)
  multipole.rs:591-856   PotentialDerivatives{1,2,3,4}::new                 (translated)
  multipole.rs:1216-1349 PotentialDerivatives::new (order 5 path)           (translated)
  multipole.rs:858-1025  gravity_{potential,accel}_multipole_oK_dK          (translated)
  multipole.rs:1352-1528 gravity_{potential,accel}_multipole (generic)      (translated)
  direct.rs:115-658      all eight direct solvers, both size branches       (translated)
  multipole.rs:1536-1595 translate_multipole (M2M)                          (loop nest re-stated here; FACT and the
                                                                             get/set index maps read from the source)
The test tests/test_oracle_vs_reference_source.py holds the oracle (and through it every GPU parity test) to these.

Run:  python tests/golden/make_reference_exec.py     (needs /root/reference; the .npz is committed)
"""
from __future__ import annotations

import math
import os
import re
import sys
from types import SimpleNamespace

import numpy as np

REF = os.environ.get("PNBX_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "crates", "gravity", "src")
HERE = os.path.dirname(os.path.abspath(__file__))
R2_TINY = sys.float_info.min  # f64::MIN_POSITIVE (multipole.rs:4, direct.rs:7)


def powi(a: float, b: int) -> float:
    """compiler-rt __powidf2 (what f64::powi lowers to): square-and-multiply."""
    recip = b < 0
    r = 1.0
    b = abs(int(b))
    while True:
        if b & 1:
            r *= a
        b //= 2
        if b == 0:
            break
        a *= a
    return 1.0 / r if recip else r


# ------------------------------------------------------------------------------------------ Rust -> Python
def strip_comment(line: str) -> str:
    return re.sub(r"//.*$", "", line).rstrip()


def expr(e: str) -> str:
    e = re.sub(r"(\d)f64\b", r"\1", e)
    e = e.replace("PotentialDerivatives::new(", "derivs_generic(")  # PotentialDerivatives4::new delegates to it
    e = re.sub(r"\b(\w+)\.len\(\)", r"len(\1)", e)
    e = re.sub(r"vec!\[\[0\.0; 3\]; (\w+)\]", r"[[0.0, 0.0, 0.0] for _ in range(\1)]", e)
    e = re.sub(r"vec!\[0\.0; (\w+)\]", r"[0.0] * \1", e)
    e = re.sub(r"\b(\w+)\.map\(\|\w+\| \w+\[(\w+)\]\)\.unwrap_or\(([\d\.]+)\)", r"(\1[\2] if \1 is not None else \3)", e)
    e = re.sub(r"\b(\w+)\.max\(([\w\.]+)\)", r"max(\1, \2)", e)
    e = e.replace(" || ", " or ").replace(" && ", " and ")
    e = re.sub(r"^\*(\w+)_i = ", r"\1[i] = ", e)  # `*acc_i = ...` inside par_iter_mut().enumerate().for_each(|(i, acc_i)|
    e = re.sub(r"\b([A-Za-z_][\w\.]*(?:\[\d+\])?)\.powi\((\w+)(?: as i32)?\)", r"powi(\1, \2)", e)
    e = re.sub(r"\b([A-Za-z_][\w\.]*)\.sqrt\(\)", r"math.sqrt(\1)", e)
    e = re.sub(r"\(([^()]*)\)\.sqrt\(\)", r"math.sqrt(\1)", e)
    assert ".sqrt()" not in e and ".powi(" not in e and " as " not in e, e
    return e


def logical_lines(lines):
    """Join rustfmt continuation lines: a statement ends with ';', a block line with '{' or '}', a field with ','."""
    buf = ""
    for raw in lines:
        s = strip_comment(raw).strip()
        if not s:
            continue
        if s.startswith("}") and buf:  # a block's tail expression (no ';') ends where the block closes
            yield buf
            buf = ""
        buf = (buf + " " + s).strip() if buf else s
        if buf.endswith((";", "{", "}", ",")) or buf.startswith("}"):
            yield buf
            buf = ""
    if buf:
        yield buf  # tail expression


def translate(lines):
    """Rust function body -> Python function body (list of lines, 4-space indents)."""
    out, ind = [], 1
    pending = list(logical_lines(lines))
    stack = []  # open Rust blocks: True = has a Python block (indent), False = braces only (`match`)
    i = 0

    def emit(s):
        out.append("    " * ind + s)

    while i < len(pending):
        s = pending[i]
        i += 1
        m = re.match(r"^(?:\} else )?if (.*) \{$", s)
        arm = re.match(r"^(\d+(?: \| \d+)*|_) => (\{\}?)$", s)
        if s.startswith("} else if "):
            ind -= 1
            emit("elif " + expr(m.group(1)) + ":")
            ind += 1
        elif m:
            emit("if " + expr(m.group(1)) + ":")
            ind += 1
            stack.append(True)
        elif s == "} else {":
            ind -= 1
            emit("else:")
            ind += 1
        elif s == "}":
            if stack.pop():
                ind -= 1
        elif re.match(r"^for (\w+) in (.+)\.\.(=?)(.+) \{$", s):
            mm = re.match(r"^for (\w+) in (.+)\.\.(=?)(.+) \{$", s)
            hi_ = mm.group(4) + (" + 1" if mm.group(3) else "")
            emit(f"for {mm.group(1)} in range({expr(mm.group(2))}, {expr(hi_)}):")
            ind += 1
            stack.append(True)
        elif re.match(r"^(\w+)\.par_iter_mut\(\)\.enumerate\(\)\.for_each\(\|\(i, \w+\)\| \{$", s):
            arr = s.split(".", 1)[0]  # rayon over targets: independent iterations, any order
            emit(f"for i in range(len({arr})):")
            ind += 1
            stack.append(True)
        elif s == "});":
            assert stack.pop()
            ind -= 1
        elif s == "let masses_slice_owned;":
            # `let masses_slice: &[f64] = if let Some(m) = masses { m } else { masses_slice_owned = vec![1.0; N]; &owned };`
            blk = []
            while pending[i] != "};":
                blk.append(pending[i])
                i += 1
            i += 1
            nvar = re.search(r"vec!\[1\.0; (\w+)\]", " ".join(blk)).group(1)
            emit(f"masses_slice = masses if masses is not None else [1.0] * {nvar}")
        elif s.startswith("for &") and s.endswith("{"):
            mm = re.match(r"^for &(\w+) in (\w+) \{$", s)
            emit(f"for {mm.group(1)} in {mm.group(2)}:")
            ind += 1
            stack.append(True)
        elif arm:  # arm of `match order.min(5) { 0 | 1 => {..} 2 => {..} _ => {..} }`
            first = not any(l.strip().startswith(("if order in", "elif order in")) for l in out)
            if arm.group(1) == "_":
                emit("else:")
            else:
                vals = ", ".join(arm.group(1).split(" | "))
                emit(("if" if first else "elif") + f" order in ({vals},):")
            ind += 1
            if arm.group(2) == "{}":
                emit("pass")
                ind -= 1
            else:
                stack.append(True)
        elif s.startswith("match order.min(5)"):
            emit("order = min(order, 5)")
            stack.append(False)
        elif s == "Self {":
            fields = []
            while pending[i] != "}":
                f = pending[i].rstrip(",")
                k, v = f.split(":", 1) if ":" in f else (f, f)  # `d100,` is shorthand for `d100: d100,`
                fields.append(f"{k.strip()}={expr(v.strip())}")
                i += 1
            i += 1
            emit("return NS(" + ", ".join(fields) + ")")
        elif s.endswith(";"):
            emit(expr(re.sub(r"^let (mut )?", "", s[:-1])))
        else:  # tail expression
            emit("return " + expr(s))
    assert not stack, stack
    return out


def function_body(path, start_pat, occurrence=0):
    """Body lines (between the item's braces) of the `occurrence`-th item whose header line matches start_pat."""
    src = open(path).read().split("\n")
    hits = [n for n, line in enumerate(src) if re.search(start_pat, line)]
    n = hits[occurrence]
    depth, body, started = 0, [], False
    for line in src[n:]:
        code = strip_comment(line)
        if not started:
            if "{" in code:
                started = True
                depth = code.count("{") - code.count("}")
                rest = code.split("{", 1)[1]
                if rest.strip():
                    body.append(rest)
            continue
        depth += code.count("{") - code.count("}")
        if depth <= 0:
            break
        body.append(line)
    return body


def make_fn(name, args, body_lines, env):
    src = f"def {name}({args}):\n" + "\n".join(body_lines) + "\n"
    exec(src, env)  # noqa: S102 - the translated reference source
    env.setdefault("__sources__", {})[name] = src
    return env[name]


def macro_bodies(path, macro):
    """{fn name: body lines} of every `macro!(name, MomentT, DerivT, m, d, { body });` invocation."""
    src = open(path).read().split("\n")
    out = {}
    n = 0
    while n < len(src):
        if src[n].startswith(macro + "!("):
            name = src[n + 1].strip().rstrip(",")
            k = n + 1
            while "{" not in strip_comment(src[k]):
                k += 1
            depth, body = 0, []
            first = strip_comment(src[k])
            depth = first.count("{") - first.count("}")
            rest = first.split("{", 1)[1]
            if depth == 0:  # one-line body: `{ expr }`
                body.append(rest.rsplit("}", 1)[0])
            else:
                if rest.strip():
                    body.append(rest)
                k += 1
                while True:
                    code = strip_comment(src[k])
                    depth += code.count("{") - code.count("}")
                    if depth <= 0:
                        break
                    body.append(src[k])
                    k += 1
            out[name] = body
            n = k
        n += 1
    return out


def build_reference_functions():
    env = {"math": math, "powi": powi, "R2_TINY": R2_TINY, "NS": SimpleNamespace, "min": min, "len": len}
    K, M = os.path.join(SRC, "kernel.rs"), os.path.join(SRC, "multipole.rs")
    make_fn("w2", "u", translate(function_body(K, r"^fn w2\(")), env)
    make_fn("w2_prime", "u", translate(function_body(K, r"^fn w2_prime\(")), env)

    text = open(M).read()
    fields = re.findall(r"pub (m\d{3}): f64", text.split("impl MultipoleMoment")[0])
    dfields = re.findall(r"pub (d\d{3}): f64", text.split("pub struct PotentialDerivatives {")[1].split("}")[0])
    assert len(fields) == 56 and len(dfields) == 56, (len(fields), len(dfields))
    env["FIELDS"], env["DFIELDS"] = fields, dfields
    env["new_moment"] = lambda: SimpleNamespace(**{f: 0.0 for f in fields})
    env["new_derivs"] = lambda: SimpleNamespace(**{f: 0.0 for f in dfields})

    # P2M (multipole.rs:82-170)
    fixed = []
    for line in function_body(M, r"fn from_points_const<"):
        line = line.replace("let mut m = MultipoleMoment::default();", "let mut m = new_moment();")
        line = line.replace("indices.is_empty()", "len(indices) == 0")
        line = line.replace("let mass = masses_opt.map(|mm| mm[pi]).unwrap_or(1.0);",
                            "let mass = masses_opt[pi] if masses_opt is not None else 1.0;")
        fixed.append(line)
    make_fn("from_points", "O, positions, masses_opt, indices, center", translate(fixed), env)

    # derivative tensors: compact orders 1-4 (multipole.rs:591-856) and the generic one (1216-1349)
    for k in range(4):
        make_fn(f"derivs{k + 1}", "dx, dy, dz, eps2",
                translate(function_body(M, r"pub fn new\(dx: f64, dy: f64, dz: f64, eps2: f64\) -> Self", k)), env)
    gen = []
    src_gen = function_body(M, r"pub fn new\(dx: f64, dy: f64, dz: f64, eps2: f64, order: u8\) -> Self")
    skip = 0
    for line in src_gen:
        if skip:
            skip -= 1
            continue
        if "let max = (order as usize).min(5);" in line:
            line = "let max = min(order, 5);"
        if "let mut d = PotentialDerivatives {" in line:  # `{ d000: dt_1, ..Default::default() };`
            gen.append("let mut d = new_derivs();")
            gen.append("d.d000 = dt_1;")
            skip = 3
            continue
        gen.append(line)
    make_fn("derivs_generic", "dx, dy, dz, eps2, order", translate(gen), env)

    # evaluators (multipole.rs:858-1025) and the generic ones (1352-1528)
    pots = macro_bodies(M, "define_multipole_potential_fn")
    accs = macro_bodies(M, "define_multipole_accel_fn")
    for name, body in {**pots, **accs}.items():
        make_fn(name, "m, d", translate(body), env)
    # direct.rs:115-658, all eight solvers (serial pair loops for N < 512 and the per-target loops alike)
    D = os.path.join(SRC, "direct.rs")
    env["kernel_potential_per_unit_mass"] = lambda kind, r, h: kernel_potential_per_unit_mass(env, kind, r, h)
    env["kernel_accel_factor"] = lambda kind, r, h: kernel_accel_factor(env, kind, r, h)
    env["max"] = max
    env["range"] = range
    for name, args in (("direct_accelerations", "positions, masses"), ("direct_accelerations_at_points", "positions, masses, targets"),
                       ("direct_potentials", "positions, masses"), ("direct_potentials_at_points", "positions, masses, targets"),
                       ("direct_potentials_kernel", "positions, masses, softenings, kernel"),
                       ("direct_accelerations_kernel", "positions, masses, softenings, kernel"),
                       ("direct_potentials_kernel_at_points", "positions, masses, softenings, targets, kernel"),
                       ("direct_accelerations_kernel_at_points", "positions, masses, softenings, targets, kernel")):
        make_fn(name, args, translate(function_body(D, rf"^pub fn {name}\(")), env)
    make_fn("gravity_potential_multipole", "m, d, order", translate(function_body(M, r"^pub fn gravity_potential_multipole\($")), env)
    make_fn("gravity_accel_multipole", "m, d, order", translate(function_body(M, r"^pub fn gravity_accel_multipole\($")), env)

    # M2M index maps and factorials from the source (multipole.rs:1027-1155)
    fact = [float(x) for x in re.search(r"const FACT: \[f64; \d+\] = \[([^\]]*)\]", text).group(1).replace("\n", " ").split(",") if x.strip()]
    getm = dict(((int(a), int(b), int(c)), f) for a, b, c, f in
                re.findall(r"\((\d), (\d), (\d)\) => m\.(m\d{3}),", text))
    assert len(getm) == 56 and fact[:6] == [1.0, 1.0, 2.0, 6.0, 24.0, 120.0]
    env["FACT"], env["LMN"] = fact, getm
    return env


def translate_multipole(env, child, shift, order):
    """multipole.rs:1536-1595, re-stated (the loop nest is control flow, not a formula): same loop order, same
    operation order per term — sign * pow / (FACT[dl] * FACT[dm] * FACT[dn]) * base, zero moments skipped."""
    FACT, LMN = env["FACT"], env["LMN"]
    o = min(int(order), 5)
    out = env["new_moment"]()
    for l in range(o + 1):
        for mm in range(o + 1):
            for n in range(o + 1):
                if l + mm + n > o:
                    continue
                s = 0.0
                for i in range(l + 1):
                    for j in range(mm + 1):
                        for k in range(n + 1):
                            base = getattr(child, LMN[(i, j, k)])
                            if base == 0.0:
                                continue
                            dl, dm, dn = l - i, mm - j, n - k
                            if dl + dm + dn == 0:
                                pw = 1.0
                            else:
                                sx = powi(shift[0], dl) if dl > 0 else 1.0
                                sy = powi(shift[1], dm) if dm > 0 else 1.0
                                sz = powi(shift[2], dn) if dn > 0 else 1.0
                                pw = sx * sy * sz
                            sign = 1.0 if (dl + dm + dn) % 2 == 0 else -1.0
                            coeff = sign * pw / (FACT[dl] * FACT[dm] * FACT[dn])
                            s += coeff * base
                setattr(out, LMN[(l, mm, n)], s)
    return out



# ------------------------------------------------------------------------------------------ tree.rs, second restatement
def fma(a, b, c):
    """f64::mul_add: exactly rounded a*b + c (rational arithmetic, then one rounding)."""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


class PyOctree:
    """tree.rs re-stated in Python, independently of oracle/gravity_oracle.cpp: the CONTROL FLOW (recursive bucket
    build :804-864, walk links :736-776, reverse payload sweeps :866-965 / :1014-1067, stackless traversal with the
    leaf-before-opening rule, zero-mass skip and the hmax softening gate :55-71 / :1069-1370, leaf sums :97-417, entry
    points :1415-1558) is written here by hand; every formula it evaluates (kernels, P2M, M2M, derivative tensors,
    evaluators) is the mechanically translated reference source. The dead "constant target softening" fast path of
    leaf_potential_sum (:122-171, unreachable through the entry points) is not restated."""
    NONE = -1

    def __init__(self, env, pos, mass, h, leaf_capacity, order, kernel):
        self.env, self.pos, self.mass, self.h = env, pos, mass, h
        self.cap, self.order_raw, self.kernel = max(int(leaf_capacity), 1), int(order), kernel
        n = len(pos)
        mn = [math.inf] * 3
        mx = [-math.inf] * 3
        for p in pos:
            for i in range(3):
                if p[i] < mn[i]:
                    mn[i] = p[i]
                if p[i] > mx[i]:
                    mx[i] = p[i]
        center = [(mn[i] + mx[i]) / 2.0 for i in range(3)]
        half = 0.0
        for i in range(3):
            half = max(half, (mx[i] - mn[i]) / 2.0)
        if half == 0.0:
            half = 1e-6
        self.nodes = [self._node(center, half, list(range(n)))]
        self._build(0)
        self._links()

    @staticmethod
    def _node(center, half, indices):
        s = half * 2.0
        return {"center": list(center), "half": half, "size2": s * s, "children": None, "indices": indices}

    def _subdivide(self, k):
        nd = self.nodes[k]
        center, half, parent = nd["center"], nd["half"], nd["indices"]
        nd["indices"] = []
        buckets = [[] for _ in range(8)]
        for pi in parent:
            p = self.pos[pi]
            oct_ = (1 if p[0] >= center[0] else 0) | (2 if p[1] >= center[1] else 0) | (4 if p[2] >= center[2] else 0)
            buckets[oct_].append(pi)
        child = [self.NONE] * 8
        for o in range(8):
            if not buckets[o]:
                continue
            cc = list(center)
            off = half / 2.0
            cc[0] += off if o & 1 else -off
            cc[1] += off if o & 2 else -off
            cc[2] += off if o & 4 else -off
            child[o] = len(self.nodes)
            self.nodes.append(self._node(cc, off, buckets[o]))
        nd["children"] = child

    def _build(self, k):
        import sys as _sys
        _sys.setrecursionlimit(10000)
        if not len(self.nodes[k]["indices"]) > self.cap:
            return
        self._subdivide(k)
        for c in self.nodes[k]["children"]:
            if c != self.NONE:
                self._build(c)

    def _links(self):
        n = len(self.nodes)
        self.first, self.next = [self.NONE] * n, [self.NONE] * n

        def rec(k):
            ch = self.nodes[k]["children"]
            if ch is None:
                return
            last = None
            for c in ch:
                if c == self.NONE:
                    continue
                if self.first[k] == self.NONE:
                    self.first[k] = c
                if last is not None:
                    self.next[last] = c
                last = c
            if last is not None:
                self.next[last] = self.next[k]
            for c in ch:
                if c != self.NONE and self.nodes[c]["children"] is not None:
                    rec(c)
        rec(0)

    def build_mass(self):
        env, nn = self.env, len(self.nodes)
        self.bh_mass, self.bh_com = [0.0] * nn, [[0.0] * 3 for _ in range(nn)]
        for k in reversed(range(nn)):
            mass, com, nd = 0.0, [0.0, 0.0, 0.0], self.nodes[k]
            if nd["children"] is None:
                if nd["indices"]:
                    for pi in nd["indices"]:
                        p = self.pos[pi]
                        if self.mass is not None:
                            m = self.mass[pi]
                            mass += m
                            com[0] += p[0] * m; com[1] += p[1] * m; com[2] += p[2] * m
                        else:
                            mass += 1.0
                            com[0] += p[0]; com[1] += p[1]; com[2] += p[2]
                    if mass > 0.0:
                        com = [c / mass for c in com]
            else:
                for c in nd["children"]:
                    if c == self.NONE or self.bh_mass[c] == 0.0:
                        continue
                    cm, cc = self.bh_mass[c], self.bh_com[c]
                    mass += cm
                    com[0] += cc[0] * cm; com[1] += cc[1] * cm; com[2] += cc[2] * cm
                if mass > 0.0:
                    com = [c / mass for c in com]
            self.bh_mass[k], self.bh_com[k] = mass, com
        self.hmax = None
        if self.h is not None:
            self.hmax = [0.0] * nn
            for k in reversed(range(nn)):
                nd, m = self.nodes[k], 0.0
                if nd["children"] is None:
                    for pi in nd["indices"]:
                        m = max(m, max(self.h[pi], 0.0))
                else:
                    for c in nd["children"]:
                        if c != self.NONE:
                            m = max(m, self.hmax[c])
                self.hmax[k] = m
        self.moments = None
        if self.order_raw > 0:
            order = min(self.order_raw, 5)
            F = env["FIELDS"]
            mom = [env["new_moment"]() for _ in range(nn)]
            for k in reversed(range(nn)):
                nd = self.nodes[k]
                if self.bh_mass[k] == 0.0:
                    continue
                if nd["children"] is None:
                    if not nd["indices"]:
                        continue
                    mom[k] = env["from_points"](order, self.pos, self.mass, nd["indices"], self.bh_com[k])
                else:
                    acc = env["new_moment"]()
                    for c in nd["children"]:
                        if c == self.NONE or self.bh_mass[c] == 0.0:
                            continue
                        shift = [self.bh_com[k][i] - self.bh_com[c][i] for i in range(3)]
                        tr = translate_multipole(env, mom[c], shift, order)
                        for f in F:  # add_assign, field by field in struct order
                            setattr(acc, f, getattr(acc, f) + getattr(tr, f))
                    mom[k] = acc
            self.moments = mom

    # ---- traversal
    def _soft_ok(self, k, dist2, th):
        if self.hmax is None:
            return True
        hh = max(self.hmax[k], 0.0)
        if th is not None:
            hh = max(hh, max(th, 0.0))
        if hh <= 0.0:
            return True
        ch = (2.8 if self.kernel == 0 else 1.0) * hh
        return dist2 > ch * ch

    def _leaf(self, indices, t, skip, th, acc_mode, out):
        env = self.env
        target_h = max(th if th is not None else 0.0, 0.0)
        use_soft = self.h is not None or target_h > 0.0
        spline = self.kernel == 1
        for pi in indices:
            if pi == skip:
                continue
            p = self.pos[pi]
            dx, dy, dz = p[0] - t[0], p[1] - t[1], p[2] - t[2]
            r2 = fma(dx, dx, fma(dy, dy, dz * dz))
            m = self.mass[pi] if self.mass is not None else 1.0
            newton = True
            hh = 0.0
            if use_soft:
                hh = max(max(self.h[pi], 0.0), target_h) if self.h is not None else target_h
                newton = hh <= 0.0 or (spline and r2 >= hh * hh)
            if not acc_mode:
                if newton:
                    inv_r = 1.0 / math.sqrt(r2 + R2_TINY)
                    out[0] += (-m * inv_r) if (use_soft or self.mass is not None) else -inv_r
                else:
                    out[0] += m * kernel_potential_per_unit_mass(env, self.kernel, math.sqrt(r2 + R2_TINY), hh)
            else:
                if newton:
                    inv_r = 1.0 / math.sqrt(r2 + R2_TINY)
                    inv_r3 = (inv_r * inv_r) * inv_r
                    if use_soft or self.mass is not None:
                        out[0] += m * dx * inv_r3; out[1] += m * dy * inv_r3; out[2] += m * dz * inv_r3
                    else:
                        out[0] += dx * inv_r3; out[1] += dy * inv_r3; out[2] += dz * inv_r3
                else:
                    g = kernel_accel_factor(env, self.kernel, math.sqrt(r2 + R2_TINY), hh)
                    out[0] += m * dx * g; out[1] += m * dy * g; out[2] += m * dz * g

    def walk(self, t, skip, th, theta, acc_mode):
        env = self.env
        out = [0.0, 0.0, 0.0] if acc_mode else [0.0]
        soft_en = self.hmax is not None or th is not None
        theta2 = theta * theta
        visits = accepts = 0
        k = 0
        while k != self.NONE:
            visits += 1
            if self.bh_mass[k] == 0.0:
                k = self.next[k]
                continue
            nd = self.nodes[k]
            if nd["children"] is None:
                self._leaf(nd["indices"], t, skip, th, acc_mode, out)
                k = self.next[k]
                continue
            com = self.bh_com[k]
            dx, dy, dz = com[0] - t[0], com[1] - t[1], com[2] - t[2]
            dist2 = fma(dx, dx, fma(dy, dy, dz * dz)) + R2_TINY
            ok = self._soft_ok(k, dist2, th) if soft_en else True
            if ok and nd["size2"] < theta2 * dist2:
                accepts += 1
                if self.moments is None:
                    inv_r = 1.0 / math.sqrt(dist2 + R2_TINY)
                    M = self.bh_mass[k]
                    if acc_mode:
                        inv_r3 = (inv_r * inv_r) * inv_r
                        out[0] += M * dx * inv_r3; out[1] += M * dy * inv_r3; out[2] += M * dz * inv_r3
                    else:
                        out[0] += -M * inv_r
                else:
                    o = min(self.order_raw, 5)
                    if o <= 1:
                        d, sfx = env["derivs1"](dx, dy, dz, R2_TINY), "o0_d1"
                    elif o <= 4:
                        d, sfx = env[f"derivs{o}"](dx, dy, dz, R2_TINY), f"o{o}_d{o}"
                    else:
                        d, sfx = env["derivs_generic"](dx, dy, dz, R2_TINY, 5), None
                    mm = self.moments[k]
                    if acc_mode:
                        a = env["gravity_accel_multipole"](mm, d, 5) if sfx is None else env["gravity_accel_multipole_" + sfx](mm, d)
                        out[0] += a[0]; out[1] += a[1]; out[2] += a[2]
                    else:
                        out[0] += env["gravity_potential_multipole"](mm, d, 5) if sfx is None else env["gravity_potential_multipole_" + sfx](mm, d)
                k = self.next[k]
            else:
                k = self.first[k]
        return out, visits, accepts

    def compute(self, theta, acc_mode):
        return [self.walk(self.pos[i], i, None if self.h is None else self.h[i], theta, acc_mode)[0] for i in range(len(self.pos))]

    def at_points(self, pts, theta, acc_mode):
        return [self.walk(p, self.NONE, None, theta, acc_mode)[0] for p in pts]


def kernel_potential_per_unit_mass(env, kind, r, h):
    """kernel.rs:41-56 (the `match kind` wrapper around w2)."""
    if r == 0.0:
        return 0.0
    if kind == 0:
        return -1.0 / math.sqrt(r * r + h * h)
    if h <= 0.0:
        return -1.0 / r
    h_inv = 1.0 / h
    u = r * h_inv
    return env["w2"](u) * h_inv


def kernel_accel_factor(env, kind, r, h):
    """kernel.rs:62-82."""
    if r == 0.0:
        return 0.0
    if kind == 0:
        s2 = r * r + h * h
        return 1.0 / (math.sqrt(s2) * s2)
    if h <= 0.0:
        return 1.0 / (r * r * r)
    h_inv = 1.0 / h
    u = r * h_inv
    return env["w2_prime"](u) * (h_inv * h_inv) / r


def evaluate(env, order, moments, dxyz, want_acc):
    """What the walk evaluates for an accepted node at `order` (dispatch of tree.rs:419-543)."""
    dx, dy, dz = dxyz
    if order <= 1:
        d, suffix = env["derivs1"](dx, dy, dz, 0.0), "o0_d1"
    elif order <= 4:
        d, suffix = env[f"derivs{order}"](dx, dy, dz, 0.0), f"o{order}_d{order}"
    else:
        d = env["derivs_generic"](dx, dy, dz, 0.0, 5)
        return (env["gravity_accel_multipole"] if want_acc else env["gravity_potential_multipole"])(moments, d, 5)
    return env[("gravity_accel_multipole_" if want_acc else "gravity_potential_multipole_") + suffix](moments, d)


def main():
    env = build_reference_functions()
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- kernels: r, h grid incl. the branch points u = 0.5, 1 and h <= 0
    r = np.concatenate([rng.uniform(1e-4, 2.0, 200), [0.25, 0.5, 0.75, 1.0, 1.5, 1e-12]])
    h = np.concatenate([rng.uniform(1e-3, 1.5, 200), [0.5, 1.0, 1.0, 1.0, 1.0, 0.3]])
    h[::17] = 0.0
    h[5::29] = -0.2
    out["kernel_r"], out["kernel_h"] = r, h
    for kind in (0, 1):
        out[f"kernel_pot_{kind}"] = np.array([kernel_potential_per_unit_mass(env, kind, a, b) for a, b in zip(r, h)])
        out[f"kernel_acc_{kind}"] = np.array([kernel_accel_factor(env, kind, a, b) for a, b in zip(r, h)])
    uu = np.concatenate([np.linspace(0.0, 1.5, 301), [0.5, 1.0, np.nextafter(0.5, 0), np.nextafter(1.0, 0)]])
    out["w2_u"] = uu
    out["w2"] = np.array([env["w2"](float(u)) for u in uu])
    out["w2_prime"] = np.array([env["w2_prime"](float(u)) if u > 0 else 0.0 for u in uu])
    # ---- P2M: 12 particles about an off-centre origin, orders 0-5, with and without masses
    pos = rng.uniform(-0.5, 0.5, (12, 3))
    mass = 0.5 + rng.random(12)
    center = pos.mean(0) + 0.01
    out["p2m_pos"], out["p2m_mass"], out["p2m_center"] = pos, mass, center
    F = env["FIELDS"]
    for O in range(6):
        m = env["from_points"](O, pos.tolist(), mass.tolist(), list(range(12)), center.tolist())
        out[f"p2m_o{O}"] = np.array([getattr(m, f) for f in F])
    m = env["from_points"](5, pos.tolist(), None, [7, 2, 9], center.tolist())
    out["p2m_unit_idx729"] = np.array([getattr(m, f) for f in F])
    # ---- M2M: translate the order-5 moments by a shift, orders 0-5
    shift = np.array([0.013, -0.021, 0.008])
    child = env["from_points"](5, pos.tolist(), mass.tolist(), list(range(12)), center.tolist())
    out["m2m_shift"] = shift
    for O in range(6):
        t = translate_multipole(env, child, shift.tolist(), O)
        out[f"m2m_o{O}"] = np.array([getattr(t, f) for f in F])
    # ---- derivative tensors and evaluators: random moments with NON-ZERO dipole (pins "no dipole in phi")
    dxyz = rng.normal(0.0, 1.0, (16, 3)) * np.array([2.0, 1.0, 0.5])
    mom = rng.normal(0.0, 1.0, (16, 56))
    mom[:, 0] = np.abs(mom[:, 0]) + 1.0
    out["m2p_dxyz"], out["m2p_moments"] = dxyz, mom
    D = env["DFIELDS"]
    out["derivs_generic_o5"] = np.array([[getattr(env["derivs_generic"](*d.tolist(), 0.0, 5), f) for f in D] for d in dxyz])
    for O in range(6):
        pot, acc = [], []
        for d, mrow in zip(dxyz, mom):
            m = SimpleNamespace(**dict(zip(F, mrow.tolist())))
            pot.append(evaluate(env, O, m, d.tolist(), False))
            acc.append(evaluate(env, O, m, d.tolist(), True))
        out[f"m2p_pot_o{O}"] = np.array(pot)
        out[f"m2p_acc_o{O}"] = np.array(acc)
    # ---- direct.rs: all eight solvers, serial pair-loop branch (N < 512) and per-target branch (N >= 512),
    # signed softenings (self mode keeps the sign in max(h_i, h_j), at-points clamps max(h_j, 0): SURVEY F14)
    for tag, n, mt in (("small", 40, 7), ("large", 520, 515)):
        pos_d = rng.uniform(-1.0, 1.0, (n, 3))
        mass_d = 0.5 + rng.random(n)
        h_d = rng.uniform(-0.05, 0.3, n)
        tgt_d = rng.uniform(-1.2, 1.2, (mt, 3))
        out[f"direct_{tag}_pos"], out[f"direct_{tag}_mass"], out[f"direct_{tag}_h"], out[f"direct_{tag}_tgt"] = pos_d, mass_d, h_d, tgt_d
        P, Mv, Hv, T = pos_d.tolist(), mass_d.tolist(), h_d.tolist(), tgt_d.tolist()
        out[f"direct_{tag}_acc"] = np.array(env["direct_accelerations"](P, Mv))
        out[f"direct_{tag}_pot"] = np.array(env["direct_potentials"](P, Mv))
        out[f"direct_{tag}_acc_unit"] = np.array(env["direct_accelerations"](P, None))
        out[f"direct_{tag}_acc_pts"] = np.array(env["direct_accelerations_at_points"](P, Mv, T))
        out[f"direct_{tag}_pot_pts"] = np.array(env["direct_potentials_at_points"](P, Mv, T))
        for kind in (0, 1):
            out[f"direct_{tag}_k{kind}_pot"] = np.array(env["direct_potentials_kernel"](P, Mv, Hv, kind))
            out[f"direct_{tag}_k{kind}_acc"] = np.array(env["direct_accelerations_kernel"](P, Mv, Hv, kind))
            out[f"direct_{tag}_k{kind}_pot_pts"] = np.array(env["direct_potentials_kernel_at_points"](P, Mv, Hv, T, kind))
            out[f"direct_{tag}_k{kind}_acc_pts"] = np.array(env["direct_accelerations_kernel_at_points"](P, Mv, Hv, T, kind))
            out[f"direct_{tag}_k{kind}_pot_noh"] = np.array(env["direct_potentials_kernel"](P, Mv, None, kind))
    # ---- tree.rs through the independent Python restatement (PyOctree): small clustered sets, every order, both kernels,
    # per-particle softenings (incl. zero / negative), unit masses, a zero-mass clump (subtree skip), theta 0.7 and 0
    def tree_case(tag, n, cap, order, kernel, with_h, with_mass, theta, zero_clump=False):
        import zlib
        r_ = np.random.default_rng(zlib.crc32(tag.encode()))
        rr = 0.3 / np.sqrt(np.maximum(r_.uniform(0, 0.99, n), 1e-9) ** (-2.0 / 3.0) - 1.0)
        v = r_.normal(size=(n, 3))
        pos_t = rr[:, None] * v / np.linalg.norm(v, axis=1)[:, None] + np.array([0.1, -0.2, 0.05])
        mass_t = (0.5 + r_.random(n)) if with_mass else None
        if zero_clump and mass_t is not None:
            sel = np.argsort(np.linalg.norm(pos_t - pos_t[0], axis=1))[:40]
            mass_t[sel] = 0.0
        h_t = r_.uniform(-0.01, 0.08, n) if with_h else None
        if h_t is not None:
            h_t[::7] = 0.0
        pts = r_.uniform(-0.6, 0.6, (9, 3))
        t = PyOctree(env, pos_t.tolist(), None if mass_t is None else mass_t.tolist(), None if h_t is None else h_t.tolist(),
                     cap, order, kernel)
        t.build_mass()
        nn = len(t.nodes)
        out[f"tree_{tag}_pos"] = pos_t
        out[f"tree_{tag}_mass"] = mass_t if mass_t is not None else np.zeros(0)
        out[f"tree_{tag}_h"] = h_t if h_t is not None else np.zeros(0)
        out[f"tree_{tag}_pts"] = pts
        out[f"tree_{tag}_params"] = np.array([cap, order, kernel, theta], dtype=np.float64)
        out[f"tree_{tag}_center"] = np.array([nd["center"] for nd in t.nodes])
        out[f"tree_{tag}_half"] = np.array([nd["half"] for nd in t.nodes])
        out[f"tree_{tag}_first"] = np.array(t.first, dtype=np.int64)
        out[f"tree_{tag}_next"] = np.array(t.next, dtype=np.int64)
        out[f"tree_{tag}_leaf_count"] = np.array([-1 if nd["children"] is not None else len(nd["indices"]) for nd in t.nodes], dtype=np.int64)
        out[f"tree_{tag}_leaf_particles"] = np.array([pi for nd in t.nodes if nd["children"] is None for pi in nd["indices"]], dtype=np.int64)
        out[f"tree_{tag}_bh_mass"] = np.array(t.bh_mass)
        out[f"tree_{tag}_bh_com"] = np.array(t.bh_com)
        if t.hmax is not None:
            out[f"tree_{tag}_hmax"] = np.array(t.hmax)
        if t.moments is not None:
            out[f"tree_{tag}_moments"] = np.array([[getattr(m_, f) for f in F] for m_ in t.moments])
        out[f"tree_{tag}_pot"] = np.array(t.compute(theta, False)).ravel()
        out[f"tree_{tag}_acc"] = np.array(t.compute(theta, True))
        out[f"tree_{tag}_pot_pts"] = np.array(t.at_points(pts.tolist(), theta, False)).ravel()
        out[f"tree_{tag}_acc_pts"] = np.array(t.at_points(pts.tolist(), theta, True))
        return nn

    cases = [("o0_plain", 260, 8, 0, 0, False, True, 0.7, False), ("o1_unit", 200, 4, 1, 0, False, False, 0.7, False),
             ("o2_spline", 300, 8, 2, 1, True, True, 0.7, False), ("o3_spline", 400, 8, 3, 1, True, True, 0.7, True),
             ("o3_plummer", 300, 6, 3, 0, True, True, 0.5, False), ("o4_spline", 240, 8, 4, 1, True, True, 0.7, False),
             ("o5_plummer", 240, 5, 5, 0, True, True, 0.7, False), ("o3_theta0", 120, 3, 3, 1, True, True, 0.0, False),
             ("o3_cap1", 150, 1, 3, 1, True, True, 0.8, False)]
    out["tree_cases"] = np.array([c[0] for c in cases])
    for c in cases:
        tree_case(*c)
    out["field_order"] = np.array(F)
    out["deriv_field_order"] = np.array(D)
    path = os.path.join(HERE, "reference_exec.npz")
    np.savez_compressed(path, **out)
    srcs = env["__sources__"]
    with open(os.path.join(HERE, "reference_exec_translated.py.txt"), "w") as f:
        f.write("# Mechanical Rust->Python translation of the reference functions executed by make_reference_exec.py\n"
                "# (kept for review; not imported by anything).\n\n")
        for name in sorted(srcs):
            f.write(srcs[name] + "\n")
    print("wrote", path, "with", len(out), "arrays; translated functions:", len(srcs))


if __name__ == "__main__":
    main()
