#!/usr/bin/env python3
"""f90fn.py -- a small mechanical Fortran 90 -> Python translator (TEST INFRASTRUCTURE).

Purpose: this image has no Fortran compiler, so the reference (100 % Fortran) cannot run here.
This tool executes the reference's own source TEXT instead: it parses the free-form Fortran of the hot-path files
(src/modm.f90, RTMmono.f90, lblrtm_sub.f90, CloudOptProp.f90, PhysConstants.f90, PlanetEarth.f90, CntnmFactors.f90,
contnm.f90, isotope.incl) statement by statement and emits Python with the same control flow, the same expression trees
(operator precedence and association preserved, integer division and real->integer truncation as Fortran defines them,
implicit typing, 1-based / custom-lower-bound arrays, argument association by reference incl. array sections and
sequence association, COMMON blocks, EQUIVALENCE of whole arrays, DATA, SAVE, OPTIONAL / keyword arguments, derived types,
SELECT CASE, labelled DO, forward GOTO).  Nothing numerical is re-derived by hand: every arithmetic statement that runs
is a token-for-token transcription of a statement in /root/reference/src.  tools/gen_ref_goldens.py drives it and writes
tests/golden/ref_*.npz; the oracle and the CUDA path are then checked against those vectors.

It is a subset translator: constructs it does not know raise TranslateError at translation time (never silently skipped),
except I/O statements (PRINT / WRITE / FORMAT / OPEN / CLOSE), which are dropped because they do not affect results.
"""
import re

from f90rt import ABSENT  # noqa: F401  (re-exported for generated code users)


class TranslateError(Exception):
    pass


# --------------------------------------------------------------------------------------------- source reading
def _strip_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def read_statements(path, include_dirs=()):
    """Logical statements of a free-form source file: [(lineno, label or None, text)], continuation lines joined,
    comments removed, INCLUDE files inlined, preprocessor lines dropped."""
    import os
    out = []
    cur, cur_ln = None, None
    with open(path, errors="replace") as f:
        raw = f.read().split("\n")
    for ln, line in enumerate(raw, 1):
        if line.startswith("#"):
            continue
        s = _strip_comment(line).rstrip()
        t = s.strip()
        if not t:
            continue
        if cur is not None:
            if t.startswith("&"):
                t = t[1:].lstrip()
            cur = cur + " " + t
        else:
            cur, cur_ln = t, ln
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        m = re.match(r"(?i)^include\s+['\"]([^'\"]+)['\"]\s*$", cur)
        if m:
            inc = None
            for d in (os.path.dirname(path),) + tuple(include_dirs):
                p = os.path.join(d, m.group(1))
                if os.path.exists(p):
                    inc = p
                    break
            if inc is None:
                raise TranslateError("%s:%d: include file %s not found" % (path, cur_ln, m.group(1)))
            out.extend(read_statements(inc, include_dirs))
        else:
            for piece in _split_semicolons(cur):
                lab = None
                m = re.match(r"^(\d+)\s+(.*)$", piece)
                if m:
                    lab, piece = int(m.group(1)), m.group(2)
                out.append((("%s:%d" % (os.path.basename(path), cur_ln)), lab, piece))
        cur = None
    return out


def _split_semicolons(s):
    if ";" not in s:
        return [s]
    parts, q, cur = [], None, ""
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            if cur.strip():
                parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


# --------------------------------------------------------------------------------------------- tokens
_TOK = re.compile(r"""
    (?P<ws>\s+)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<num>(?:\d+(?:\.(?![a-z]{2,5}\.)\d*)?|\.\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)
  | (?P<id>[a-z_]\w*)
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|[-+*/(),:=<>%])
""", re.X)


def tokenize(text, where=""):
    toks = []
    i = 0
    low = _lower_outside_strings(text)
    while i < len(low):
        m = _TOK.match(low, i)
        if not m:
            raise TranslateError("%s: cannot tokenize at %r" % (where, low[i:i + 30]))
        i = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        toks.append((k, m.group(k)))
    return toks


def _lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
            out.append(ch.lower())
    return "".join(out)


# --------------------------------------------------------------------------------------------- expression parser
class P:
    """Recursive-descent parser over a token list.  AST nodes are tuples:
    ('num', text) ('str', s) ('log', bool) ('name', id) ('app', id, [args]) ('bin', op, l, r) ('un', op, x)
    ('mem', base, field) ('kw', name, expr) ('sl', lo, hi, step) ('cx', re, im)"""

    def __init__(self, toks, where=""):
        self.t, self.i, self.where = toks, 0, where

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise TranslateError("%s: expected %r, got %r in %s" % (self.where, val, self.peek()[1], self.text()))

    def text(self):
        return " ".join(v for _, v in self.t)

    def done(self):
        return self.i >= len(self.t)

    # precedence, lowest first
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        l = self.p_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            l = ("bin", op, l, self.p_or())
        return l

    def p_or(self):
        l = self.p_and()
        while self.peek()[1] == ".or.":
            self.next()
            l = ("bin", ".or.", l, self.p_and())
        return l

    def p_and(self):
        l = self.p_not()
        while self.peek()[1] == ".and.":
            self.next()
            l = ("bin", ".and.", l, self.p_not())
        return l

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.next()
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    _REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=",
            "==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}

    def p_rel(self):
        l = self.p_cat()
        if self.peek()[1] in self._REL:
            op = self._REL[self.next()[1]]
            return ("bin", op, l, self.p_cat())
        return l

    def p_cat(self):
        l = self.p_add()
        while self.peek()[1] == "//":
            self.next()
            l = ("bin", "//", l, self.p_add())
        return l

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            l = ("un", op, self.p_mul())
        else:
            l = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            l = ("bin", op, l, self.p_mul())
        return l

    def p_mul(self):
        l = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            l = ("bin", op, l, self.p_pow())
        return l

    def p_pow(self):
        l = self.p_prim()
        if self.peek()[1] == "**":
            self.next()
            # right associative; the exponent may carry a unary sign (a**-b is accepted by gfortran)
            if self.peek()[1] in ("+", "-"):
                op = self.next()[1]
                r = ("un", op, self.p_pow())
            else:
                r = self.p_pow()
            return ("bin", "**", l, r)
        return l

    def p_prim(self):
        k, v = self.next()
        if k == "num":
            return ("num", v)
        if k == "str":
            q = v[0]
            return ("str", v[1:-1].replace(q + q, q))
        if k == "dotop" and v in (".true.", ".false."):
            return ("log", v == ".true.")
        if v == "(":
            e = self.expr()
            if self.accept(","):
                im = self.expr()
                self.expect(")")
                return ("cx", e, im)
            self.expect(")")
            return ("par", e)
        if k == "id":
            node = ("name", v)
            if self.peek()[1] == "(":
                self.next()
                args = self.arglist()
                node = ("app", v, args)
            while self.peek()[1] == "%":
                self.next()
                fk, fv = self.next()
                if fk != "id":
                    raise TranslateError("%s: bad component reference" % self.where)
                node = ("mem", node, fv)
                if self.peek()[1] == "(":
                    raise TranslateError("%s: subscripted component not supported" % self.where)
            return node
        raise TranslateError("%s: unexpected token %r in %s" % (self.where, v, self.text()))

    def arglist(self):
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.arg())
            if self.accept(")"):
                return args
            self.expect(",")

    def arg(self):
        # keyword argument?
        if self.peek()[0] == "id" and self.peek(1)[1] == "=" and self.peek(2)[1] != "=":
            name = self.next()[1]
            self.next()
            return ("kw", name, self.expr())
        lo = hi = st = None
        if self.peek()[1] != ":":
            lo = self.expr()
            if self.peek()[1] != ":":
                return lo
        self.expect(":")
        if self.peek()[1] not in (",", ")", ":"):
            hi = self.expr()
        if self.accept(":"):
            st = self.expr()
        return ("sl", lo, hi, st)


def parse_expr(text, where=""):
    p = P(tokenize(text, where), where)
    e = p.expr()
    if not p.done():
        raise TranslateError("%s: trailing tokens in expression %r" % (where, text))
    return e


# --------------------------------------------------------------------------------------------- program structure
INTRINSICS = {
    "abs", "dabs", "iabs", "cabs", "exp", "dexp", "cexp", "log", "alog", "dlog", "log10", "alog10", "sqrt", "dsqrt", "tanh",
    "sin", "cos", "tan", "atan", "asin", "acos", "sinh", "cosh", "atan2", "int", "ifix", "idint", "nint", "anint", "aint",
    "real", "float", "dble", "sngl", "aimag", "imag", "cmplx", "dcmplx", "conjg", "max", "min", "amax1", "amin1", "max0",
    "min0", "dmax1", "dmin1", "mod", "amod", "dmod", "sign", "isign", "dsign", "sum", "maxval", "minval", "size", "present",
    "trim", "len_trim", "dim", "epsilon", "tiny", "huge",
}
PY_RESERVED = {"and", "as", "assert", "async", "await", "break", "class", "continue", "def", "del", "elif", "else", "except",
               "finally", "for", "from", "global", "if", "import", "in", "is", "lambda", "nonlocal", "not", "or", "pass",
               "raise", "return", "try", "while", "with", "yield", "rt", "np", "None", "True", "False", "int", "float",
               "complex", "abs", "max", "min", "sum", "len", "range", "print", "type", "id", "str", "all", "any"}


def py(name):
    return name + "_" if name in PY_RESERVED else name


class Var:
    def __init__(self, name):
        self.name = name
        self.type = None          # 'int' 'real' 'complex' 'logical' 'char' 'type:<name>'
        self.dims = None          # list of (lo_ast or None, hi_ast or None / '*' / ':') or None for scalars
        self.is_dummy = False
        self.optional = False
        self.param = None         # AST of a PARAMETER value
        self.init = None          # AST of an initialisation expression (module variable / SAVEd local)
        self.data = None          # DATA items in source order: dict(kind='whole'|'implied', vals=[ASTs], subs, loopvar, lo, hi)
        self.common = None        # (block, index)
        self.saved = False
        self.equiv = None         # name of the variable whose storage this one shares (whole-array EQUIVALENCE)
        self.use = None           # (module, remote name) for USE association


class Unit:
    """A procedure (subroutine / function), a module, or a block data unit."""

    def __init__(self, kind, name, args=None, result=None, parent=None):
        self.kind, self.name, self.args, self.result, self.parent = kind, name, args or [], result, parent
        self.vars = {}
        self.implicit = {c: ("int" if c in "ijklmn" else "real") for c in "abcdefghijklmnopqrstuvwxyz"}
        self.implicit_none = False
        self.stmts = []            # executable statements: (where, label, text)
        self.children = []         # module procedures
        self.types = {}            # derived types defined here: name -> [fields]
        self.generic = {}          # INTERFACE name -> specific procedure
        self.uses = []             # (module, only list or None)
        self.modified = set()      # dummy names this procedure assigns (directly or through callees)
        self.prefix_type = None
        self.has_goto = False
        self.body = None

    def var(self, name):
        v = self.vars.get(name)
        if v is None:
            v = self.vars[name] = Var(name)
        return v


_TYPE_RE = re.compile(r"^(integer|real|double\s*precision|complex|logical|character|type\s*\(\s*(\w+)\s*\))"
                      r"(\s*\*\s*(\d+|\(\s*\*\s*\))|\s*\(\s*(?:len\s*=\s*|kind\s*=\s*)?[\w*]+\s*\))?")


def _split_top(s, sep=","):
    parts, depth, q, cur = [], 0, None, ""
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
            cur += ch
        elif ch in "([":
            depth += 1
            cur += ch
        elif ch in ")]":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip() or parts:
        parts.append(cur.strip())
    return parts


def _find_unquoted(s, ch, start):
    q = None
    for i in range(start, len(s)):
        c = s[i]
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == ch:
            return i
    raise ValueError("no unquoted %r" % ch)


def _parse_dims(text, where):
    dims = []
    for d in _split_top(text):
        d = d.strip()
        if d == "*":
            dims.append((None, "*"))
        elif d == ":":
            dims.append((None, ":"))
        elif ":" in d and _split_top(d, ":") and len(_split_top(d, ":")) == 2:
            lo, hi = _split_top(d, ":")
            dims.append((parse_expr(lo, where) if lo else None, "*" if hi == "*" else (parse_expr(hi, where) if hi else ":")))
        else:
            dims.append((None, parse_expr(d, where)))
    return dims


class Translator:
    def __init__(self):
        self.units = {}        # procedures by name (lower case)
        self.modules = {}
        self.blockdata = []
        self.commons = {}      # block -> list of (rank signature) from the first declaration
        self.externals = {}    # name -> list of modified dummy positions (python callables supplied at load time)
        self.order = []

    # ------------------------------------------------------------------ pass 1: structure + declarations
    def add_file(self, path, include_dirs=()):
        stmts = read_statements(path, include_dirs)
        i = 0
        stack = []         # open units
        cur = None
        in_interface = None
        in_type = None
        contains_mode = False
        for where, label, text in stmts:
            low = _lower_outside_strings(text).strip()
            # --- unit boundaries
            m = re.match(r"^module\s+(\w+)\s*$", low)
            if m and not low.startswith("module procedure"):
                cur = Unit("module", m.group(1))
                self.modules[cur.name] = cur
                stack.append(cur)
                continue
            m = re.match(r"^block\s*data(?:\s+(\w+))?\s*$", low)
            if m:
                cur = Unit("blockdata", m.group(1) or "blockdata%d" % len(self.blockdata))
                self.blockdata.append(cur)
                stack.append(cur)
                continue
            m = re.match(r"^program\s+(\w+)", low)
            if m:
                cur = Unit("program", m.group(1))
                stack.append(cur)
                continue
            if in_interface is not None:
                m = re.match(r"^module\s+procedure\s+(\w+)", low)
                if m and in_interface:
                    stack[-1].generic[in_interface] = m.group(1)
                if re.match(r"^end\s*interface", low):
                    in_interface = None
                continue
            m = re.match(r"^interface(?:\s+(\w+))?\s*$", low)
            if m:
                in_interface = m.group(1) or ""
                continue
            if in_type is not None:
                if re.match(r"^end\s*type", low):
                    in_type = None
                    continue
                mm = _TYPE_RE.match(low)
                if not mm:
                    raise TranslateError("%s: unsupported statement in TYPE: %s" % (where, text))
                rest = low[mm.end():]
                rest = rest.split("::", 1)[1] if "::" in rest else rest
                for ent in _split_top(rest):
                    in_type[1].append(re.match(r"\w+", ent.strip()).group(0))
                continue
            m = re.match(r"^type\s+(\w+)\s*$", low) or re.match(r"^type\s*::\s*(\w+)\s*$", low)
            if m:
                in_type = (m.group(1), [])
                stack[-1].types[m.group(1)] = in_type[1]
                continue
            m = re.match(r"^(?:(recursive|pure|elemental)\s+)?(?:(integer|real(?:\s*\*\s*\d+)?|double\s*precision|complex|logical)\s+)?"
                         r"(subroutine|function)\s+(\w+)\s*(?:\(([^)]*)\))?\s*(?:result\s*\(\s*(\w+)\s*\))?\s*$", low)
            if m:
                kind, name = m.group(3), m.group(4)
                args = [a.strip() for a in (m.group(5) or "").split(",") if a.strip()]
                parent = stack[-1] if stack and stack[-1].kind == "module" else None
                u = Unit(kind, name, args, m.group(6) or (name if kind == "function" else None), parent)
                u.prefix_type = m.group(2)
                for a in args:
                    u.var(a).is_dummy = True
                if u.prefix_type and u.result:
                    u.var(u.result).type = self._type_name(u.prefix_type)
                self.units[name] = u
                self.order.append(name)
                if parent:
                    parent.children.append(u)
                stack.append(u)
                cur = u
                continue
            if low == "contains":
                continue
            if re.match(r"^end\s*(subroutine|function|module|program|block\s*data)?(\s+\w+)?\s*$", low) and \
                    not re.match(r"^end\s*(if|do|select|type|interface|where)", low):
                if not stack:
                    raise TranslateError("%s: END without an open unit" % where)
                stack.pop()
                cur = stack[-1] if stack else None
                continue
            if cur is None:
                raise TranslateError("%s: statement outside a program unit: %s" % (where, text))
            if not self._declaration(cur, where, low, text):
                cur.stmts.append((where, label, text))

    @staticmethod
    def _type_name(t):
        t = re.sub(r"\s+", "", t)
        if t.startswith("integer"):
            return "int"
        if t.startswith("real") or t.startswith("doubleprecision"):
            return "real"
        if t.startswith("complex"):
            return "complex"
        if t.startswith("logical"):
            return "logical"
        if t.startswith("character"):
            return "char"
        m = re.match(r"type\((\w+)\)", t)
        if m:
            return "type:" + m.group(1)
        raise TranslateError("unknown type " + t)

    def _declaration(self, u, where, low, text):
        """Returns True when the statement is a declaration (and records it)."""
        if low.startswith("implicit"):
            if "none" in low:
                u.implicit_none = True
                return True
            for m in re.finditer(r"(integer|real\s*\*\s*\d+|real|double\s*precision|complex|logical|character\s*\*\s*\d+)\s*\(([^)]*)\)", low[8:]):
                tn = self._type_name(m.group(1))
                for rng in m.group(2).split(","):
                    rng = rng.strip()
                    a, b = (rng.split("-") + [rng])[:2] if "-" in rng else (rng, rng)
                    for c in range(ord(a.strip()), ord(b.strip()) + 1):
                        u.implicit[chr(c)] = tn
            return True
        if re.match(r"^(private|public|save\s*$|external|intrinsic|intent|optional\b)", low):
            if low.startswith("optional"):
                for n in _split_top(low[8:].replace("::", "")):
                    u.var(n.strip()).optional = True
            return True
        if low.startswith("save"):
            for n in _split_top(low[4:].replace("::", "")):
                n = n.strip().strip("/")
                if n:
                    u.var(n).saved = True
            return True
        m = re.match(r"^use\s+(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", low)
        if m:
            only = None
            if m.group(2) is not None:
                only = []
                for ent in _split_top(m.group(2)):
                    if "=>" in ent:
                        loc, rem = [x.strip() for x in ent.split("=>")]
                    else:
                        loc = rem = ent.strip()
                    if loc:
                        only.append((loc, rem))
            u.uses.append((m.group(1), only))
            return True
        m = re.match(r"^parameter\s*\((.*)\)\s*$", low)
        if m:
            for ent in _split_top(self._orig_tail(text, "(")[:-1] if False else m.group(1)):
                k, v = ent.split("=", 1)
                u.var(k.strip()).param = parse_expr(v.strip(), where)
            return True
        m = re.match(r"^dimension\s+(.*)$", low)
        if m:
            for ent in _split_top(m.group(1)):
                mm = re.match(r"^(\w+)\s*\((.*)\)$", ent.strip())
                u.var(mm.group(1)).dims = _parse_dims(mm.group(2), where)
            return True
        m = re.match(r"^common\s*/\s*(\w+)\s*/\s*(.*)$", low)
        if m:
            blk = m.group(1)
            members = []
            for ent in _split_top(m.group(2)):
                mm = re.match(r"^(\w+)\s*(?:\((.*)\))?$", ent.strip())
                v = u.var(mm.group(1))
                if mm.group(2) is not None:
                    v.dims = _parse_dims(mm.group(2), where)
                members.append(v)
            for idx, v in enumerate(members):
                v.common = (blk, idx)
            self.commons.setdefault(blk, []).append((u, members))
            return True
        if low.startswith("equivalence"):
            for grp in re.findall(r"\(([^()]*(?:\([^()]*\)[^()]*)*)\)", low[len("equivalence"):]):
                names = [n.strip() for n in _split_top(grp)]
                if any("(" in n for n in names):
                    # element equivalence (e.g. an integer flag overlaid on a real array element): the alias reads 0
                    for n in names:
                        if "(" not in n:
                            u.var(n).equiv = "@zero"
                    continue
                for n in names[1:]:
                    u.var(n).equiv = names[0]
            return True
        if low.startswith("data") and re.match(r"^data\s*[\w(]", low):
            self._data(u, where, text)
            return True
        mm = _TYPE_RE.match(low)
        if mm and not re.match(r"^(real|integer|complex|logical)\s*\(", low[:mm.end()] + "x") or (mm and "::" in low):
            return self._type_decl(u, where, low, text, mm)
        if mm:
            # 'real(...)' could be an assignment to an array named real -- not in this code base
            return self._type_decl(u, where, low, text, mm)
        if re.match(r"^(format|namelist)\b", low):
            return True
        return False

    @staticmethod
    def _orig_tail(text, ch):
        return text[text.index(ch) + 1:]

    def _type_decl(self, u, where, low, text, mm):
        tn = self._type_name(mm.group(1))
        rest = low[mm.end():]
        # keep the original-case text for initialisers (string literals)
        attrs = ""
        if "::" in rest:
            attrs, rest = rest.split("::", 1)
        elif rest.lstrip().startswith(","):
            raise TranslateError("%s: attribute list without '::'" % where)
        if re.match(r"^\s*function\b", rest):
            return False
        attr_list = [a.strip() for a in _split_top(attrs) if a.strip()]
        dim_attr = None
        is_param = optional = saved = False
        for a in attr_list:
            if a.startswith("dimension"):
                dim_attr = _parse_dims(a[a.index("(") + 1:a.rindex(")")], where)
            elif a == "parameter":
                is_param = True
            elif a == "optional":
                optional = True
            elif a == "save":
                saved = True
            elif a.startswith("intent") or a in ("target", "pointer", "allocatable", "private", "public", "external"):
                pass
            else:
                raise TranslateError("%s: unsupported attribute %r" % (where, a))
        for ent in _split_top(rest):
            ent = ent.strip()
            if not ent:
                continue
            init = None
            if "=" in ent and not re.match(r"^\w+\s*\([^)]*=", ent):
                ent, init = ent.split("=", 1)
                ent = ent.strip()
            m2 = re.match(r"^(\w+)\s*(?:\((.*)\))?\s*(?:\*\s*\d+)?$", ent)
            if not m2:
                raise TranslateError("%s: cannot parse entity %r" % (where, ent))
            v = u.var(m2.group(1))
            v.type = tn
            if m2.group(2) is not None:
                v.dims = _parse_dims(m2.group(2), where)
            elif dim_attr is not None:
                v.dims = dim_attr
            v.optional = v.optional or optional
            v.saved = v.saved or saved
            if init is not None:
                e = parse_expr(init.strip(), where)
                if is_param:
                    v.param = e
                else:
                    v.init = e
                    v.saved = True
        return True

    def _data(self, u, where, text):
        body = _lower_outside_strings(text)[4:].strip()
        # split into  <object list> / <values> /  groups
        pos = 0
        while pos < len(body):
            m = re.match(r"\s*,?\s*", body[pos:])
            pos += m.end()
            if pos >= len(body):
                break
            s1 = _find_unquoted(body, "/", pos)
            s2 = _find_unquoted(body, "/", s1 + 1)
            objs, vals = body[pos:s1].strip(), body[s1 + 1:s2]
            pos = s2 + 1
            values = []
            for tok in _split_top(vals):
                tok = tok.strip()
                if not tok:
                    continue
                rep = 1
                mm = re.match(r"^(\d+)\s*\*\s*(.+)$", tok)
                if mm:
                    rep, tok = int(mm.group(1)), mm.group(2)
                values.extend([parse_expr(tok, where)] * rep)
            objlist = _split_top(objs)
            # implied DO over one array:  (x(subscripts), i = lo, hi)
            if len(objlist) == 1 and objlist[0].startswith("("):
                inner = objlist[0][1:-1]
                parts = _split_top(inner)
                if len(parts) != 3:
                    raise TranslateError("%s: unsupported implied-DO in DATA: %s" % (where, objlist[0]))
                obj = parse_expr(parts[0], where)
                lv, lo = [x.strip() for x in parts[1].split("=")]
                if obj[0] != "app":
                    raise TranslateError("%s: unsupported implied-DO object: %s" % (where, parts[0]))
                v = u.var(obj[1])
                v.data = (v.data or []) + [dict(kind="implied", vals=values, subs=obj[2], loopvar=lv,
                                                lo=parse_expr(lo, where), hi=parse_expr(parts[2], where))]
                v.saved = True
                continue
            if len(objlist) == 1:
                v = u.var(objlist[0])
                v.data = (v.data or []) + [dict(kind="whole", vals=values)]
                v.saved = True
            else:
                if len(objlist) != len(values):
                    raise TranslateError("%s: DATA object/value count mismatch" % where)
                for o, val in zip(objlist, values):
                    v = u.var(o)
                    v.data = (v.data or []) + [dict(kind="whole", vals=[val])]
                    v.saved = True

    # ------------------------------------------------------------------ COMMON storage association
    def const_eval(self, u, e):
        """value of a constant expression (literals, PARAMETERs, + - * /, parentheses) at translation time"""
        k = e[0]
        if k == "num":
            t = re.sub(r"_\w+$", "", e[1])
            return int(t) if re.match(r"^\d+$", t) else float(t.replace("d", "e"))
        if k == "par":
            return self.const_eval(u, e[1])
        if k == "un":
            v = self.const_eval(u, e[2])
            return -v if e[1] == "-" else v
        if k == "name":
            r = self.resolve(u, e[1])
            if r is None or r[0].param is None:
                raise TranslateError("%s: %s is not a constant" % (u.name, e[1]))
            v = self.const_eval(r[1], r[0].param)
            return int(v) if self.var_type(r[1], e[1]) == "int" else v
        if k == "bin":
            a, b = self.const_eval(u, e[2]), self.const_eval(u, e[3])
            if e[1] == "+":
                return a + b
            if e[1] == "-":
                return a - b
            if e[1] == "*":
                return a * b
            if e[1] == "/":
                if isinstance(a, int) and isinstance(b, int):
                    q = abs(a) // abs(b)
                    return q if (a >= 0) == (b >= 0) else -q
                return a / b
            if e[1] == "**":
                return a ** b
        raise TranslateError("%s: cannot evaluate constant expression %r" % (u.name, e))

    def const_shape(self, u, v):
        shape, lbs = [], []
        for lo, hi in v.dims:
            l = int(self.const_eval(u, lo)) if lo is not None else 1
            if hi in ("*", ":"):
                raise TranslateError("%s: COMMON array %s needs constant bounds" % (u.name, v.name))
            h = int(self.const_eval(u, hi))
            shape.append(h - l + 1)
            lbs.append(l)
        return shape, lbs

    def layout_commons(self):
        """storage association: every member gets the offset of its first storage unit (all numeric scalars are one
        8-byte unit in the parity build; CHARACTER*8 items count as one unit too)"""
        self.common_size = {}
        for blk, decls in self.commons.items():
            total = 0
            for u, members in decls:
                off = 0
                for v in members:
                    n = 1
                    if v.dims is not None:
                        shape, lbs = self.const_shape(u, v)
                        for x in shape:
                            n *= x
                        v.common_shape, v.common_lbs = shape, lbs
                    v.common_off, v.common_n = off, n
                    off += n
                total = max(total, off)
            self.common_size[blk] = total

    # ------------------------------------------------------------------ symbol resolution
    def resolve(self, u, name):
        """-> (Var, owner) following USE association and host (module) association; None if unknown."""
        if name in u.vars:
            v = u.vars[name]
            if v.use is None:
                return v, u
        for mod, only in u.uses:
            mu = self.modules.get(mod)
            if mu is None:
                continue
            if only is None:
                if name in mu.vars:
                    return mu.vars[name], mu
            else:
                for loc, rem in only:
                    if loc == name and rem in mu.vars:
                        return mu.vars[rem], mu
        if u.parent is not None:
            r = self.resolve(u.parent, name)
            if r:
                return r
        return None

    def var_type(self, u, name):
        r = self.resolve(u, name)
        if r and r[0].type:
            return r[0].type
        if r and r[0].param is not None and r[0].type is None:
            pass
        owner = r[1] if r else u
        return owner.implicit.get(name[0], "real")

    def is_array(self, u, name):
        r = self.resolve(u, name)
        return bool(r and r[0].dims is not None)

    def find_proc(self, u, name):
        """procedure (Unit) reachable from u under this name, or None"""
        seen = name
        for scope in (u, u.parent):
            if scope is None:
                continue
            if seen in scope.generic:
                seen = scope.generic[seen]
            for mod, only in scope.uses:
                mu = self.modules.get(mod)
                if mu is None:
                    continue
                cand = seen
                if only is not None:
                    hit = [rem for loc, rem in only if loc == seen]
                    if not hit:
                        continue
                    cand = hit[0]
                cand = mu.generic.get(cand, cand)
                for c in mu.children:
                    if c.name == cand:
                        return c
        return self.units.get(seen)

    # ------------------------------------------------------------------ pass 2: statements -> block tree
    def build_body(self, u):
        if u.body is not None:
            return
        items = []
        for where, label, text in u.stmts:
            items.append(self._classify(u, where, label, text))
        self._pos = 0
        self._dostack = []
        self._items = items
        u.body = self._block(u, terminators=())
        if self._pos != len(items):
            raise TranslateError("%s: unbalanced block structure near %s" % (u.name, items[self._pos][1]))
        labels = {}
        self._collect_labels(u.body, labels)
        u.labels = labels

    def _collect_labels(self, block, labels):
        for st in block:
            if st.get("label") is not None:
                labels[st["label"]] = st
            for key in ("body", "orelse"):
                if key in st and st[key]:
                    self._collect_labels(st[key], labels)
            for br in st.get("branches", []):
                self._collect_labels(br[1], labels)

    def _classify(self, u, where, label, text):
        low = _lower_outside_strings(text).strip()
        d = dict(where=where, label=label, text=text, low=low)
        if re.match(r"^(print\b|write\s*\(|read\s*\(|open\s*\(|close\s*\(|format\s*\(|rewind\b|backspace\b|inquire\s*\()", low):
            d["k"] = "io"
        elif re.match(r"^if\s*\(", low):
            cond, rest = self._paren_split(low[2:].strip(), where)
            if rest == "then":
                d.update(k="if_then", cond=cond)
            else:
                inner = self._classify(u, where, None, self._tail_orig(text, rest))
                d.update(k="if_stmt", cond=cond, inner=inner)
        elif re.match(r"^else\s*if\s*\(", low):
            cond, rest = self._paren_split(low[low.index("("):], where)
            d.update(k="elseif", cond=cond)
        elif low == "else":
            d["k"] = "else"
        elif re.match(r"^end\s*if$", low):
            d["k"] = "endif"
        elif re.match(r"^do\s+while\s*\(", low):
            cond, rest = self._paren_split(low[low.index("("):], where)
            d.update(k="do_while", cond=cond)
        elif re.match(r"^do\s+(\d+\s+)?\w+\s*=", low):
            m = re.match(r"^do\s+(?:(\d+)\s+)?(\w+)\s*=\s*(.*)$", low)
            parts = _split_top(m.group(3))
            d.update(k="do", dolabel=int(m.group(1)) if m.group(1) else None, var=m.group(2), lim=parts)
        elif low == "do":
            raise TranslateError("%s: bare DO not supported" % where)
        elif re.match(r"^end\s*do$", low):
            d["k"] = "enddo"
        elif low == "continue":
            d["k"] = "continue"
        elif re.match(r"^go\s*to\s+\d+$", low):
            d.update(k="goto", target=int(re.search(r"\d+", low).group(0)))
            u.has_goto = True
        elif re.match(r"^go\s*to\s*\(", low):
            raise TranslateError("%s: computed GOTO not supported" % where)
        elif re.match(r"^call\s+\w+", low):
            m = re.match(r"^call\s+(\w+)\s*(\(.*\))?\s*$", low)
            args = []
            if m.group(2):
                p = P(tokenize(self._tail_orig(text, m.group(2)), where), where)
                p.expect("(")
                args = p.arglist()
            d.update(k="call", name=m.group(1), args=args)
        elif low == "return":
            d["k"] = "return"
        elif re.match(r"^stop\b", low):
            d.update(k="stop", msg=text[4:].strip().strip("'\""))
        elif re.match(r"^select\s*case\s*\(", low):
            cond, rest = self._paren_split(low[low.index("("):], where)
            d.update(k="select", cond=cond)
        elif re.match(r"^case\s*\(", low):
            inner = low[low.index("(") + 1:low.rindex(")")]
            d.update(k="case", items=_split_top(inner))
        elif re.match(r"^case\s+default$", low):
            d.update(k="case", items=None)
        elif re.match(r"^end\s*select$", low):
            d["k"] = "endselect"
        elif low in ("exit", "cycle"):
            d["k"] = low
        else:
            # assignment
            toks = tokenize(text, where)
            p = P(toks, where)
            lhs = p.p_prim()
            if not p.accept("="):
                raise TranslateError("%s: unsupported statement: %s" % (where, text))
            rhs = p.expr()
            if not p.done():
                raise TranslateError("%s: trailing tokens in assignment: %s" % (where, text))
            d.update(k="assign", lhs=lhs, rhs=rhs)
        return d

    @staticmethod
    def _tail_orig(text, low_tail):
        """original-case tail of `text` that corresponds to the lower-cased tail `low_tail`"""
        return text[len(text) - len(low_tail):] if low_tail else ""

    def _paren_split(self, s, where):
        """s starts with '(' : returns (AST of the parenthesised expression, remaining text stripped)"""
        depth, q = 0, None
        for i, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    return parse_expr(s[1:i], where), s[i + 1:].strip()
        raise TranslateError("%s: unbalanced parentheses" % where)

    def _block(self, u, terminators, dolabel=None):
        out = []
        while self._pos < len(self._items):
            st = self._items[self._pos]
            k = st["k"]
            if k in terminators:
                return out
            if dolabel is not None and st.get("label") == dolabel:
                # the terminating statement of a labelled DO belongs to the loop body; nested loops may share it
                shared = self._dostack.count(dolabel) > 1
                if shared and k not in ("continue", "enddo"):
                    raise TranslateError("%s: nested DO loops share an executable terminal statement" % st["where"])
                if not shared:
                    self._pos += 1
                if k == "enddo":
                    st = dict(st, k="continue")
                out.append(self._nest(u, st))
                return out
            self._pos += 1
            out.append(self._nest(u, st))
        if terminators or dolabel is not None:
            raise TranslateError("%s: block not terminated (%s)" % (u.name, terminators or dolabel))
        return out

    def _nest(self, u, st):
        k = st["k"]
        if k == "if_then":
            branches = []
            cond = st["cond"]
            while True:
                body = self._block(u, ("elseif", "else", "endif"))
                branches.append((cond, body))
                t = self._items[self._pos]
                self._pos += 1
                if t["k"] == "endif":
                    break
                cond = t["cond"] if t["k"] == "elseif" else None
            return dict(st, k="if", branches=branches)
        if k == "do":
            if st["dolabel"] is not None:
                self._dostack.append(st["dolabel"])
                body = self._block(u, (), dolabel=st["dolabel"])
                self._dostack.pop()
            else:
                body = self._block(u, ("enddo",))
                self._pos += 1
            return dict(st, body=body)
        if k == "do_while":
            body = self._block(u, ("enddo",))
            self._pos += 1
            return dict(st, body=body)
        if k == "select":
            cases = []
            first = self._block(u, ("case", "endselect"))
            if first:
                raise TranslateError("%s: statements before the first CASE" % st["where"])
            while True:
                t = self._items[self._pos]
                self._pos += 1
                if t["k"] == "endselect":
                    break
                body = self._block(u, ("case", "endselect"))
                cases.append((t["items"], body))
            return dict(st, cases=cases)
        return st

    # ------------------------------------------------------------------ pass 3: which dummies does a procedure modify
    def analyse_modified(self):
        for u in self.units.values():
            self.build_body(u)
        changed = True
        while changed:
            changed = False
            for u in self.units.values():
                before = len(u.modified)
                self._scan_modified(u, u.body)
                if len(u.modified) != before:
                    changed = True

    def _scan_modified(self, u, block):
        for st in block:
            k = st["k"]
            if k == "assign":
                base = self._base_name(st["lhs"])
                if base in u.args:
                    u.modified.add(base)
                self._scan_expr_calls(u, st["rhs"])
            elif k == "call":
                self._scan_call(u, st["name"], st["args"])
            elif k == "if_stmt":
                self._scan_modified(u, [st["inner"]])
                self._scan_expr_calls(u, st["cond"])
            elif k == "if":
                for cond, body in st["branches"]:
                    if cond is not None:
                        self._scan_expr_calls(u, cond)
                    self._scan_modified(u, body)
            elif k in ("do", "do_while"):
                if k == "do" and st["var"] in u.args:
                    u.modified.add(st["var"])
                self._scan_modified(u, st["body"])
            elif k == "select":
                for items, body in st["cases"]:
                    self._scan_modified(u, body)

    def _scan_call(self, u, name, args):
        callee = self.find_proc(u, name)
        if callee is None:
            return
        for pos, a in enumerate(args):
            dn = a[1] if a[0] == "kw" else (callee.args[pos] if pos < len(callee.args) else None)
            ae = a[2] if a[0] == "kw" else a
            if dn in callee.modified:
                base = self._base_name(ae)
                if base in u.args:
                    u.modified.add(base)

    def _scan_expr_calls(self, u, e):
        if not isinstance(e, tuple):
            return
        if e[0] == "app" and not self.is_array(u, e[1]) and e[1] not in INTRINSICS:
            self._scan_call(u, e[1], e[2])
        for x in e[1:]:
            if isinstance(x, tuple):
                self._scan_expr_calls(u, x)
            elif isinstance(x, list):
                for y in x:
                    self._scan_expr_calls(u, y)

    @staticmethod
    def _base_name(e):
        while isinstance(e, tuple) and e[0] in ("mem", "par"):
            e = e[1]
        if isinstance(e, tuple) and e[0] in ("name", "app"):
            return e[1]
        return None

    # ------------------------------------------------------------------ pass 4: emission
    def emit_module(self, names=None):
        """Python source for every translated unit (or the closure of `names`)."""
        self.analyse_modified()
        self.layout_commons()
        out = ["# generated by tools/f90fn.py from the reference's Fortran source -- do not edit", "import numpy as np",
               "import f90rt as rt", ""]
        # derived types
        for mod in self.modules.values():
            for tname, fields in mod.types.items():
                out.append("class T_%s:" % tname)
                out.append("    __slots__ = (%s)" % "".join("%r, " % f for f in fields))
                out.append("    def __init__(self, %s):" % ", ".join("%s=0.0" % py(f) for f in fields))
                for f in fields:
                    out.append("        self.%s = %s" % (f, py(f)))
                out.append("    def _copy(self):")
                out.append("        return T_%s(%s)" % (tname, ", ".join("self.%s" % f for f in fields)))
                out.append("")
        # COMMON blocks
        for blk, decls in self.commons.items():
            out.append("C_%s = rt.Namespace()" % blk)
            out.append("C_%s.s = np.zeros(%d, dtype=object)" % (blk, self.common_size[blk]))
            out.append("C_%s.s[:] = 0.0" % blk)
        out.append("")
        # module variables
        for mod in self.modules.values():
            out.append("M_%s = rt.Namespace()" % mod.name)
        em = Emitter(self)
        for mod in self.modules.values():
            out.extend(em.module_init(mod))
        out.append("")
        for name in self.order:
            u = self.units[name]
            out.extend(em.procedure(u))
            out.append("")
        # COMMON arrays are allocated by the first routine that dimensions them; BLOCK DATA fills DATA values
        for bd in self.blockdata:
            out.extend(em.blockdata(bd))
        out.extend(em.common_init())
        return "\n".join(out) + "\n"


class Emitter:
    def __init__(self, tr):
        self.tr = tr
        self.tmp = 0

    # ---------------------------------------------------------------- names
    def ref(self, u, name, as_value=False):
        """Python expression for variable `name` seen from unit u."""
        r = self.tr.resolve(u, name)
        if r is None:
            v, owner = u.var(name), u
        else:
            v, owner = r
        if v.equiv == "@zero":
            return "0"
        if v.equiv:
            return self.ref(owner if owner is u else owner, v.equiv)
        if v.common is not None:
            if v.dims is None:
                return "C_%s.s[%d]" % (v.common[0], v.common_off)
            if owner is u and u.kind in ("subroutine", "function"):
                return py(v.name)            # bound to a view of the block's storage in the prologue
            return self.common_view(v)
        if owner.kind == "module":
            return "M_%s.%s" % (owner.name, v.name)
        if v.param is not None and not v.is_dummy:
            return py(v.name)
        if (v.saved or v.data is not None or v.init is not None) and not v.is_dummy:
            return "S_%s.%s" % (u.name, v.name)
        return py(v.name)

    @staticmethod
    def common_view(v):
        return "rt.FArr(C_%s.s[%d:%d].reshape((%s,), order='F'), (%s,))" % (
            v.common[0], v.common_off, v.common_off + v.common_n, ", ".join(str(x) for x in v.common_shape),
            ", ".join(str(x) for x in v.common_lbs))

    # ---------------------------------------------------------------- expressions
    def expr(self, u, e, pre, want_array_value=True):
        k = e[0]
        if k == "num":
            return self.num(e[1])
        if k == "str":
            return repr(e[1])
        if k == "log":
            return "True" if e[1] else "False"
        if k == "par":
            return "(" + self.expr(u, e[1], pre) + ")"
        if k == "cx":
            return "complex(%s, %s)" % (self.expr(u, e[1], pre), self.expr(u, e[2], pre))
        if k == "name":
            r = self.ref(u, e[1])
            if self.tr.is_array(u, e[1]) and want_array_value:
                return "rt.val(%s)" % r
            return r
        if k == "mem":
            return "%s.%s" % (self.expr(u, e[1], pre, False), e[2])
        if k == "un":
            x = self.expr(u, e[2], pre)
            if e[1] == ".not.":
                return "(not %s)" % x
            return "(%s%s)" % (e[1], x)
        if k == "bin":
            op = e[1]
            l, r = self.expr(u, e[2], pre), self.expr(u, e[3], pre)
            if op == "/":
                return "rt.div(%s, %s)" % (l, r)
            if op == "**":
                return "rt.pw(%s, %s)" % (l, r)
            if op == ".and.":
                return "(%s and %s)" % (l, r)
            if op == ".or.":
                return "(%s or %s)" % (l, r)
            if op == ".eqv.":
                return "(bool(%s) == bool(%s))" % (l, r)
            if op == ".neqv.":
                return "(bool(%s) != bool(%s))" % (l, r)
            if op == "//":
                return "rt.concat(%s, %s)" % (l, r)
            return "(%s %s %s)" % (l, op, r)
        if k == "app":
            return self.app(u, e, pre)
        raise TranslateError("cannot emit %r" % (e,))

    @staticmethod
    def num(t):
        t = re.sub(r"_\w+$", "", t)
        if re.match(r"^\d+$", t):
            return str(int(t))
        t = t.replace("d", "e")
        if t.endswith("."):
            t += "0"
        if t.startswith("."):
            t = "0" + t
        t = re.sub(r"\.e", ".0e", t)
        return repr(float(t)) if True else t

    def subscripts(self, u, args, pre):
        parts = []
        for a in args:
            if a[0] == "sl":
                lo = self.expr(u, a[1], pre) if a[1] is not None else ""
                hi = self.expr(u, a[2], pre) if a[2] is not None else ""
                s = "%s:%s" % (lo, hi)
                if a[3] is not None:
                    s += ":" + self.expr(u, a[3], pre)
                parts.append(s)
            else:
                parts.append(self.expr(u, a, pre))
        return ", ".join(parts)

    def app(self, u, e, pre):
        name, args = e[1], e[2]
        if self.tr.is_array(u, name):
            return "%s[%s]" % (self.ref(u, name), self.subscripts(u, args, pre))
        r = self.tr.resolve(u, name)
        if r and r[0].type == "char" and r[0].dims is None:
            raise TranslateError("character substring not supported: %s" % name)
        # structure constructor
        for mod in self.tr.modules.values():
            if name in mod.types:
                return "T_%s(%s)" % (name, ", ".join(self.expr(u, a, pre) for a in args))
        callee = self.tr.find_proc(u, name)
        if callee is None and name in INTRINSICS:
            return "rt.f_%s(%s)" % (name, ", ".join(self.callarg_plain(u, a, pre) for a in args))
        if callee is None and name in self.tr.externals:
            return "%s(%s)[0]" % (name, ", ".join(self.callarg_plain(u, a, pre) for a in args))
        if callee is None:
            raise TranslateError("%s: unknown function or array %r" % (u.name, name))
        call, writeback = self.call(u, callee, args, pre)
        if writeback:
            self.tmp += 1
            t = "_r%d" % self.tmp
            pre.append("%s = %s" % (t, call))
            pre.extend(w.replace("@R", t) for w in writeback)
            return "%s[0]" % t
        return "%s[0]" % call

    def callarg_plain(self, u, a, pre):
        if a[0] == "kw":
            return "%s=%s" % (py(a[1]), self.expr(u, a[2], pre))
        return self.expr(u, a, pre)

    def call(self, u, callee, args, pre):
        """-> (python call expression, [write-back statements using @R for the result tuple])"""
        parts, wb = [], []
        mod_list = [a for a in callee.args if a in callee.modified and callee.vars[a].dims is None]
        for pos, a in enumerate(args):
            if a[0] == "kw":
                dn, ae = a[1], a[2]
            else:
                if pos >= len(callee.args):
                    raise TranslateError("%s: too many arguments in call to %s" % (u.name, callee.name))
                dn, ae = callee.args[pos], a
            dv = callee.vars.get(dn)
            dummy_is_array = bool(dv and dv.dims is not None)
            if dummy_is_array:
                src = self.array_actual(u, ae, pre)
            else:
                src = self.expr(u, ae, pre, want_array_value=False)
            parts.append("%s=%s" % (py(dn), src) if a[0] == "kw" else src)
            if dn in callee.modified and not dummy_is_array:
                tgt = self.lvalue(u, ae, pre)
                if tgt is not None:
                    wb.append(self.store(u, ae, tgt, "@R[%d]" % (1 + mod_list.index(dn))))
        return "%s(%s)" % (py(callee.name), ", ".join(parts)), wb

    def array_actual(self, u, ae, pre):
        """actual argument bound to an array dummy: whole array, section (view) or element (sequence association)"""
        if ae[0] == "name":
            return self.ref(u, ae[1])
        if ae[0] == "app" and self.tr.is_array(u, ae[1]):
            if any(x[0] == "sl" for x in ae[2]):
                return "%s[%s]" % (self.ref(u, ae[1]), self.subscripts(u, ae[2], pre))
            return "%s.flat_from(%s)" % (self.ref(u, ae[1]), self.subscripts(u, ae[2], pre))
        raise TranslateError("%s: unsupported actual for an array dummy: %r" % (u.name, ae))

    def lvalue(self, u, e, pre):
        """Python assignment target for an assignable expression, else None"""
        while e[0] == "par":
            return None
        if e[0] == "name":
            r = self.tr.resolve(u, e[1])
            if r and r[0].param is not None:
                return None
            return self.ref(u, e[1])
        if e[0] == "app" and self.tr.is_array(u, e[1]):
            return "%s[%s]" % (self.ref(u, e[1]), self.subscripts(u, e[2], pre))
        if e[0] == "mem":
            return "%s.%s" % (self.expr(u, e[1], pre, False), e[2])
        return None

    def store(self, u, lhs, target, value):
        """assignment statement with the conversion Fortran applies for the target's type"""
        base = self.tr._base_name(lhs)
        if lhs[0] == "name":
            if self.tr.is_array(u, base):
                return "%s.assign(%s)" % (target, value)
            t = self.tr.var_type(u, base)
            if t == "int":
                return "%s = rt.toint(%s)" % (target, value)
            if t == "real":
                return "%s = rt.toreal(%s)" % (target, value)
            if t == "complex":
                return "%s = rt.tocomplex(%s)" % (target, value)
            if t.startswith("type:"):
                return "%s = rt.copyval(%s)" % (target, value)
            return "%s = %s" % (target, value)
        return "%s = %s" % (target, value)

    # ---------------------------------------------------------------- statements
    def block(self, u, stmts, ind, loop_labels=()):
        # a label of this statement list that a LATER goto names (a backward jump: the hand-written loops "1000 continue ...
        # goto 1000"): the list runs inside "while True"; a pending jump to one of its labels restarts the list, and the
        # guards skip everything up to the label
        back = sorted(st["label"] for st in stmts if st.get("label") is not None and u.has_goto and
                      st["label"] in getattr(u, "back_targets", ()))
        if back:
            if any(st["k"] in ("exit", "cycle") for st in stmts):
                raise TranslateError("%s: EXIT/CYCLE next to the target of a backward GOTO" % u.name)
            inner = self._block_plain(u, stmts, ind + "    ")
            return [ind + "while True:"] + inner + ["%s    if _g in (%s,): continue" % (ind, ", ".join(str(b) for b in back)),
                                                    ind + "    break"]
        return self._block_plain(u, stmts, ind)

    def _block_plain(self, u, stmts, ind):
        out = []
        g = u.has_goto
        for st in stmts:
            if st.get("label") is not None and g and st["label"] in u.goto_targets:
                out.append("%sif _g == %d: _g = 0" % (ind, st["label"]))
            out.extend(self.stmt(u, st, ind))
        if not out:
            out.append(ind + "pass")
        return out

    def guard(self, u, ind, lines):
        if not u.has_goto:
            return [ind + l for l in lines]
        return [ind + "if _g == 0:"] + [ind + "    " + l for l in lines]

    def stmt(self, u, st, ind):
        k = st["k"]
        g = u.has_goto
        pre = []
        if k in ("io", "continue"):
            return []
        if k == "assign":
            rhs = self.expr(u, st["rhs"], pre)
            tgt = self.lvalue(u, st["lhs"], pre)
            if tgt is None:
                raise TranslateError("%s: not assignable: %s" % (st["where"], st["text"]))
            return self.guard(u, ind, pre + [self.store(u, st["lhs"], tgt, rhs)])
        if k == "call":
            return self.guard(u, ind, self.call_stmt(u, st))
        if k == "goto":
            return self.guard(u, ind, ["_g = %d" % st["target"]])
        if k == "return":
            return self.guard(u, ind, ["return " + self.ret(u)])
        if k == "stop":
            return self.guard(u, ind, ["raise rt.FortranStop(%r)" % st["msg"]])
        if k == "if_stmt":
            cond = self.expr(u, st["cond"], pre)
            inner = self.stmt(dict_no_goto(u), st["inner"], "") or ["pass"]
            lines = pre + ["if %s:" % cond] + ["    " + l for l in inner]
            return self.guard(u, ind, lines)
        if k == "if":
            lines = []
            first = True
            for cond, body in st["branches"]:
                if cond is not None:
                    pc = []
                    c = self.expr(u, cond, pc)
                    if pc and not first:
                        raise TranslateError("%s: side-effecting call in ELSE IF condition" % st["where"])
                    lines.extend(ind + p_ for p_ in pc)
                    kw = "if" if first else "elif"
                    lines.append("%s%s %s%s:" % (ind, kw, "_g == 0 and " if g else "", c))
                else:
                    lines.append("%s%s" % (ind, "elif _g == 0:" if g else "else:"))
                lines.extend(self.block(u, body, ind + "    "))
                first = False
            return lines
        if k == "do":
            lim = [self.expr(u, parse_expr(x, st["where"]), pre) for x in st["lim"]]
            var = self.ref(u, st["var"])
            lines = [ind + p_ for p_ in pre]
            ind2 = ind
            if g:
                lines.append(ind + "if _g == 0:")
                ind2 = ind + "    "
            lines.append("%sfor %s in rt.do(%s):" % (ind2, var, ", ".join(lim)))
            body = self.block(u, st["body"], ind2 + "    ")
            lines.extend(body)
            if g:
                if st["dolabel"] is not None and st["dolabel"] in u.goto_targets and \
                        not (st["body"] and st["body"][-1].get("label") == st["dolabel"]):
                    lines.append("%s    if _g == %d: _g = 0" % (ind2, st["dolabel"]))
                lines.append(ind2 + "    if _g != 0: break")
            lines.append(ind2 + "else:")
            lines.append("%s    %s = rt.do_final(%s)" % (ind2, var, ", ".join(lim)))
            return lines
        if k == "do_while":
            cond = self.expr(u, st["cond"], pre)
            if pre:
                raise TranslateError("%s: side-effecting call in DO WHILE condition" % st["where"])
            lines = []
            ind2 = ind
            if g:
                lines.append(ind + "if _g == 0:")
                ind2 = ind + "    "
            lines.append("%swhile %s:" % (ind2, cond))
            lines.extend(self.block(u, st["body"], ind2 + "    "))
            if g:
                lines.append(ind2 + "    if _g != 0: break")
            return lines
        if k == "select":
            sel = self.expr(u, st["cond"], pre)
            self.tmp += 1
            sv = "_s%d" % self.tmp
            lines = [ind + p_ for p_ in pre] + ["%s%s = %s" % (ind, sv, sel)]
            first = True
            default = None
            for items, body in st["cases"]:
                if items is None:
                    default = body
                    continue
                conds = []
                for it in items:
                    if ":" in it:
                        lo, hi = [x.strip() for x in it.split(":")]
                        conds.append("rt.inrange(%s, %s, %s)" % (sv, self.expr(u, parse_expr(lo), pre) if lo else "None",
                                                                 self.expr(u, parse_expr(hi), pre) if hi else "None"))
                    else:
                        conds.append("%s == %s" % (sv, self.expr(u, parse_expr(it), pre)))
                lines.append("%s%s %s(%s):" % (ind, "if" if first else "elif", "_g == 0 and " if g else "", " or ".join(conds)))
                lines.extend(self.block(u, body, ind + "    "))
                first = False
            if default is not None:
                if first:
                    lines.append(ind + "if True:")
                else:
                    lines.append(ind + ("elif _g == 0:" if g else "else:"))
                lines.extend(self.block(u, default, ind + "    "))
            return lines
        if k == "exit":
            return self.guard(u, ind, ["break"])
        if k == "cycle":
            return self.guard(u, ind, ["continue"])
        raise TranslateError("%s: cannot emit statement kind %s" % (st["where"], k))

    def call_stmt(self, u, st):
        pre = []
        name = st["name"]
        callee = self.tr.find_proc(u, name)
        if callee is None:
            if name in self.tr.externals:
                args = ", ".join(self.ext_arg(u, a, pre) for a in st["args"])
                return pre + ["%s(%s)" % (name, args)]
            raise TranslateError("%s: call to unknown procedure %s" % (st["where"], name))
        call, wb = self.call(u, callee, st["args"], pre)
        if wb:
            return pre + ["_r = " + call] + [w.replace("@R", "_r") for w in wb]
        return pre + [call]

    def ext_arg(self, u, a, pre):
        if a[0] == "kw":
            return "%s=%s" % (py(a[1]), self.ext_arg(u, a[2], pre))
        if a[0] == "name" and self.tr.is_array(u, a[1]):
            return self.ref(u, a[1])
        if a[0] == "app" and self.tr.is_array(u, a[1]) and any(x[0] == "sl" for x in a[2]):
            return "%s[%s]" % (self.ref(u, a[1]), self.subscripts(u, a[2], pre))
        return self.expr(u, a, pre, want_array_value=False)

    def ret(self, u):
        vals = [self.ref(u, u.result) if u.result else "None"]
        for a in u.args:
            if a in u.modified and u.vars[a].dims is None:       # arrays are associated by reference
                vals.append(self.ref(u, a))
        return "(" + ", ".join(vals) + ("," if len(vals) == 1 else "") + ")"

    def conv(self, u, name, val):
        t = self.tr.var_type(u, name)
        return {"int": "rt.toint(%s)", "real": "rt.toreal(%s)", "complex": "rt.tocomplex(%s)"}.get(t, "%s") % val

    # ---------------------------------------------------------------- procedures
    def dims_src(self, u, v, pre):
        """-> (shape source list or None when assumed, lower-bound source list)"""
        shape, lbs = [], []
        assumed = False
        for lo, hi in v.dims:
            los = self.expr(u, lo, pre) if lo is not None else "1"
            lbs.append(los)
            if hi in ("*", ":"):
                assumed = True
                shape.append(None)
            else:
                his = self.expr(u, hi, pre)
                shape.append("(%s) - (%s) + 1" % (his, los) if lo is not None else his)
        return (None if assumed else shape), lbs, shape

    _DT = {"int": "np.int64", "real": "np.float64", "complex": "np.complex128", "logical": "np.bool_"}

    def procedure(self, u):
        tr = self.tr
        u.goto_targets = set()
        self._find_targets(u, u.body, u.goto_targets)
        u.back_targets = set()
        self._find_back_targets(u.body, set(), u.back_targets)
        sig = []
        for a in u.args:
            v = u.vars[a]
            sig.append("%s=rt.ABSENT" % py(a) if v.optional else py(a))
        out = ["def %s(%s):" % (py(u.name), ", ".join(sig))]
        ind = "    "
        body = []
        # parameters first (they may size arrays), in declaration order
        for v in u.vars.values():
            if v.param is not None and not v.is_dummy:
                pre = []
                body.append("%s%s = %s" % (ind, py(v.name), self.conv(u, v.name, self.expr(u, v.param, pre))))
        # dummy arrays: rebind with the declared rank / lower bounds
        for a in u.args:
            v = u.vars[a]
            if v.dims is not None:
                pre = []
                _, lbs, _ = self.dims_src(u, v, pre)
                body.append("%s%s = rt.dummy(%s, %d, (%s,))" % (ind, py(a), py(a), len(v.dims), ", ".join(lbs)))
        # local arrays (automatic / explicit-shape, not saved, not common, not equivalenced)
        for v in u.vars.values():
            if v.dims is None or v.is_dummy or v.common is not None or v.use is not None or v.param is not None:
                continue
            if v.equiv:
                continue
            if v.saved or v.data is not None:
                continue
            pre = []
            shape, lbs, _ = self.dims_src(u, v, pre)
            if shape is None:
                raise TranslateError("%s: assumed-size local array %s" % (u.name, v.name))
            dt = self._DT.get(tr.var_type(u, v.name), "object")
            body.append("%s%s = rt.FArr.zeros((%s,), (%s,), %s)" % (ind, py(v.name), ", ".join(shape), ", ".join(lbs), dt))
        # COMMON arrays: views of the block's storage sequence
        for v in u.vars.values():
            if v.common is not None and v.dims is not None:
                body.append("%s%s = %s" % (ind, py(v.name), self.common_view(v)))
        # plain local scalars start at zero (undefined in Fortran; the reference never relies on a value)
        names = set()
        self._collect_names(u.body, names)
        for n in sorted(names):
            r = tr.resolve(u, n)
            if r is not None and r[1] is not u:
                continue
            v = u.vars.get(n)
            if v is not None and (v.is_dummy or v.dims is not None or v.common is not None or v.param is not None or v.equiv or
                                  v.saved or v.data is not None or v.init is not None or v.use is not None):
                continue
            if n == u.result:
                continue
            if v is None and (tr.find_proc(u, n) is not None or n in INTRINSICS or n in tr.externals):
                continue
            t = tr.var_type(u, n)
            if t.startswith("type:"):
                body.append("%s%s = T_%s()" % (ind, py(n), t[5:]))
            else:
                body.append("%s%s = %s" % (ind, py(n), {"int": "0", "complex": "0j", "logical": "False", "char": "''"}.get(t, "0.0")))
        if u.result and u.result not in u.args:
            t = tr.var_type(u, u.result)
            body.append("%s%s = %s" % (ind, py(u.result), {"int": "0", "complex": "0j"}.get(t, "0.0")))
        if u.has_goto:
            body.append(ind + "_g = 0")
        body.extend(self.block(u, u.body, ind))
        if u.has_goto:
            body.append(ind + "if _g != 0: raise rt.FortranStop('GOTO %d: label not reached (backward or inward jump) in " + u.name + "' % _g)")
        body.append(ind + "return " + self.ret(u))
        out.extend(body)
        # statics (SAVE / DATA / initialised) of this procedure
        st = ["S_%s = rt.Namespace()" % u.name]
        for v in u.vars.values():
            if v.is_dummy or v.common is not None or v.param is not None or v.equiv:
                continue
            if not (v.saved or v.data is not None or v.init is not None):
                continue
            st.extend(self.static_init(u, v, "S_%s.%s" % (u.name, v.name)))
        return st + out

    def data_fill(self, u, v, target, pre):
        lines = []
        for it in v.data or []:
            vals = "[%s]" % ", ".join(self.expr(u, x, pre) for x in it["vals"])
            if it["kind"] == "implied":
                subs = []
                for sub in it["subs"]:
                    if sub == ("name", it["loopvar"]):
                        subs.append("%s:%s" % (self.expr(u, it["lo"], pre), self.expr(u, it["hi"], pre)))
                    else:
                        subs.append(self.expr(u, sub, pre))
                lines.append("%s[%s] = %s" % (target, ", ".join(subs), vals))
            else:
                lines.append("_v = %s" % vals)
                lines.append("%s.a.reshape(-1, order='F')[:len(_v)] = _v" % target)
        return lines

    def static_init(self, u, v, target):
        tr = self.tr
        pre = []
        lines = []
        t = tr.var_type(u, v.name)
        if v.dims is not None:
            shape, lbs, _ = self.dims_src(u, v, pre)
            if shape is None:
                raise TranslateError("%s: assumed-shape static %s" % (u.name, v.name))
            dt = self._DT.get(t, "object")
            lines.append("%s = rt.FArr.zeros((%s,), (%s,), %s)" % (target, ", ".join(shape), ", ".join(lbs), dt))
            lines.extend(self.data_fill(u, v, target, pre))
        else:
            if v.data is not None:
                val = self.expr(u, v.data[0]["vals"][0], pre)
            elif v.init is not None:
                val = self.expr(u, v.init, pre)
            else:
                val = {"int": "0", "complex": "0j", "logical": "False", "char": "''"}.get(t, "0.0")
            conv = {"int": "rt.toint(%s)", "real": "rt.toreal(%s)", "complex": "rt.tocomplex(%s)"}.get(t, "%s")
            lines.append("%s = %s" % (target, conv % val))
        # PARAMETERs used in the dimension / data expressions must be visible at module level
        need = []
        for pv in u.vars.values():
            if pv.param is not None and not pv.is_dummy:
                need.append("%s = %s" % (py(pv.name), self.conv(u, pv.name, self.expr(u, pv.param, []))))
        return need + lines

    def _collect_names(self, block, acc):
        def walk(e):
            if isinstance(e, tuple):
                if e and e[0] == "name":
                    acc.add(e[1])
                elif e and e[0] == "app":
                    acc.add(e[1])
                    for a in e[2]:
                        walk(a)
                elif e and e[0] == "kw":
                    walk(e[2])
                else:
                    for x in e[1:]:
                        walk(x)
            elif isinstance(e, list):
                for x in e:
                    walk(x)
        for st in block:
            for key in ("lhs", "rhs", "cond"):
                if key in st:
                    walk(st[key])
            if st["k"] == "call":
                walk(st["args"])
            if st["k"] == "do":
                acc.add(st["var"])
                for x in st["lim"]:
                    walk(parse_expr(x))
            if st["k"] == "if_stmt":
                self._collect_names([st["inner"]], acc)
            if "body" in st:
                self._collect_names(st["body"], acc)
            for br in st.get("branches", []):
                if br[0] is not None:
                    walk(br[0])
                self._collect_names(br[1], acc)
            for cs in st.get("cases", []):
                self._collect_names(cs[1], acc)

    def _find_back_targets(self, block, seen, acc):
        """labels that a goto names after they have been passed in textual order"""
        for st in block:
            if st.get("label") is not None:
                seen.add(st["label"])
            tgt = st["target"] if st["k"] == "goto" else (st["inner"]["target"] if st["k"] == "if_stmt" and st["inner"]["k"] == "goto" else None)
            if tgt is not None and tgt in seen:
                acc.add(tgt)
            if "body" in st:
                self._find_back_targets(st["body"], seen, acc)
            for br in st.get("branches", []):
                self._find_back_targets(br[1], seen, acc)
            for cs in st.get("cases", []):
                self._find_back_targets(cs[1], seen, acc)

    def _find_targets(self, u, block, acc):
        for st in block:
            if st["k"] == "goto":
                acc.add(st["target"])
            if st["k"] == "if_stmt" and st["inner"]["k"] == "goto":
                acc.add(st["inner"]["target"])
            for key in ("body",):
                if key in st:
                    self._find_targets(u, st[key], acc)
            for br in st.get("branches", []):
                self._find_targets(u, br[1], acc)
            for cs in st.get("cases", []):
                self._find_targets(u, cs[1], acc)

    def module_init(self, mod):
        out = []
        for v in mod.vars.values():
            if v.use is not None:
                continue
            pre = []
            tgt = "M_%s.%s" % (mod.name, v.name)
            if v.common is not None:
                out.extend(self.common_member_init(mod, v))
                continue
            if v.param is not None:
                val = self.expr(mod, v.param, pre)
                t = self.tr.var_type(mod, v.name)
                if t.startswith("type:"):
                    out.append("%s = None  # filled after the type constructors exist" % tgt)
                    out.append("%s = %s" % (tgt, val))
                else:
                    out.append("%s = %s" % (tgt, self.conv(mod, v.name, val)))
            elif v.dims is not None:
                shape, lbs, _ = self.dims_src(mod, v, pre)
                if shape is None:
                    out.append("%s = None" % tgt)
                    continue
                dt = self._DT.get(self.tr.var_type(mod, v.name), "object")
                out.append("%s = rt.FArr.zeros((%s,), (%s,), %s)" % (tgt, ", ".join(shape), ", ".join(lbs), dt))
                out.extend(self.data_fill(mod, v, tgt, pre))
            elif v.init is not None:
                out.append("%s = %s" % (tgt, self.conv(mod, v.name, self.expr(mod, v.init, pre))))
            elif v.data is not None:
                out.append("%s = %s" % (tgt, self.conv(mod, v.name, self.expr(mod, v.data[0]["vals"][0], pre))))
            else:
                t = self.tr.var_type(mod, v.name)
                out.append("%s = %s" % (tgt, {"int": "0", "complex": "0j", "logical": "False", "char": "''"}.get(t, "0.0")))
        return out

    def common_member_init(self, u, v):
        out = []
        pre = []
        if v.dims is not None:
            out.extend(self.data_fill(u, v, self.common_view(v), pre))
        elif v.data is not None:
            out.append("C_%s.s[%d] = %s" % (v.common[0], v.common_off, self.conv(u, v.name, self.expr(u, v.data[0]["vals"][0], pre))))
        return out

    def blockdata(self, bd):
        out = ["# BLOCK DATA %s" % bd.name]
        for v in bd.vars.values():
            if v.param is not None:
                out.append("%s = %s" % (py(v.name), self.conv(bd, v.name, self.expr(bd, v.param, []))))
        for v in bd.vars.values():
            if v.common is not None:
                out.extend(self.common_member_init(bd, v))
        return out

    def common_init(self):
        """integer scalar members of COMMON blocks that nothing initialised start as integer zero"""
        out = []
        for blk, decls in self.tr.commons.items():
            for u, members in decls:
                for v in members:
                    if v.dims is None and self.tr.var_type(u, v.name) == "int":
                        out.append("if isinstance(C_%s.s[%d], float) and C_%s.s[%d] == 0.0: C_%s.s[%d] = 0" %
                                   (blk, v.common_off, blk, v.common_off, blk, v.common_off))
        return out


class dict_no_goto:
    """view of a unit with has_goto switched off (the inner statement of a logical IF is guarded by the IF itself)"""

    def __init__(self, u):
        self.__dict__["_u"] = u

    def __getattr__(self, k):
        if k == "has_goto":
            return False
        return getattr(self._u, k)


def load(paths, include_dirs=(), externals=None, extra_globals=None, dump=None):
    """Translate the given source files and execute the result; returns the namespace (dict).
    externals: {name: python callable} for procedures that are not translated (they receive the actual arguments and
    must return a tuple whose first element is the function value)."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    tr = Translator()
    tr.externals = dict(externals or {})
    for p in paths:
        tr.add_file(p, include_dirs)
    src = tr.emit_module()
    if dump:
        with open(dump, "w") as f:
            f.write(src)
    ns = dict(extra_globals or {})
    ns.update(tr.externals)
    exec(compile(src, dump or "<f90fn>", "exec"), ns)
    ns["__translator__"] = tr
    return ns


if __name__ == "__main__":
    import sys
    tr = Translator()
    for p in sys.argv[1:]:
        tr.add_file(p)
    sys.stdout.write(tr.emit_module())
