#!/usr/bin/env python3
"""Extract the numeric DATA tables the hot path needs from the reference Fortran
and emit them as C initialiser lists under monortm_b200/csrc/tables/.

Only NUMBERS are taken (scientific coefficient tables); no reference code is
copied.  Every literal is parsed with Python's correctly-rounded float() -- the
same decimal->binary64 conversion gfortran performs for default-real literals
under -fdefault-real-8 (build/makefile.common:195-198) -- and written back with
repr(), which round-trips the binary64 value exactly.

Sources (relative to /root/reference/src):
  contnm.f90:186-202     XFAC_RHU(-1:61)    H2O foreign MW/far-IR scale factors
  contnm.f90:1473-1936   BLOCK DATA BS296   H2O self continuum 296 K  (2003 pts)
  contnm.f90:1981-2444   BLOCK DATA BS260   H2O self continuum 260 K  (2003 pts)
  contnm.f90:2489-2954   BLOCK DATA BFH2O   H2O foreign continuum     (2003 pts)
  contnm.f90:3018-4156   BLOCK DATA BFCO2   CO2 continuum             (5003 pts)
  contnm.f90:2969-2975   tdep_bandhead(1196:1220)
  contnm.f90:4232-4327   BN2T296 / BN2T220  N2 roto-translational CIA (73 pts x4)
  tips_2003.f90          QofT(iso,1:119) for all 38 tabulated molecules, ISONM, Tdat
  isotope.incl           SMASS(39,9) isotopologue masses

Run here (the reference is not present on the GPU box); outputs are committed.
"""
import os
import re
import sys

REF = os.environ.get("MONORTM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "monortm_b200", "csrc", "tables")

NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eEdD][-+]?\d+)?"


def read_lines(path):
    with open(path, "r", errors="replace") as f:
        return f.read().split("\n")


def strip_comment(line):
    # none of the table units contain '!' inside character literals
    i = line.find("!")
    return line if i < 0 else line[:i]


def join_statements(lines):
    """Free-form Fortran: join '&' continuations, drop comments."""
    stmts, cur = [], ""
    for raw in lines:
        s = strip_comment(raw).rstrip()
        if not s.strip():
            continue
        t = s.strip()
        if t.startswith("&"):
            t = t[1:].lstrip()
        cont = t.endswith("&")
        if cont:
            t = t[:-1].rstrip()
        cur += " " + t
        if not cont:
            stmts.append(cur.strip())
            cur = ""
    if cur.strip():
        stmts.append(cur.strip())
    return stmts


def parse_values(body):
    vals = []
    for tok in body.split(","):
        tok = tok.strip()
        if not tok:
            continue
        rep = 1
        if "*" in tok:
            r, tok = tok.split("*")
            rep = int(r)
        tok = tok.strip().replace("d", "e").replace("D", "e")
        if tok.endswith("."):
            tok += "0"
        vals.extend([float(tok)] * rep)
    return vals


def data_statements(stmts):
    """Yield (name, [values]) for every 'DATA name / ... /' (one name per stmt)."""
    for s in stmts:
        m = re.match(r"(?i)^data\s+(.+?)\s*/(.*)/\s*$", s)
        if not m:
            continue
        yield m.group(1).strip(), m.group(2)


def block(lines, start_pat, end_pat):
    i0 = next(i for i, l in enumerate(lines) if re.match(start_pat, l.strip(), re.I))
    i1 = next(i for i in range(i0 + 1, len(lines)) if re.match(end_pat, lines[i].strip(), re.I))
    return lines[i0:i1 + 1], i0 + 1, i1 + 1


def emit(f, ctype, name, vals, cite, per_line=5):
    f.write("/* %s */\n" % cite)
    f.write("static const %s %s[%d] = {\n" % (ctype, name, len(vals)))
    for i in range(0, len(vals), per_line):
        chunk = vals[i:i + per_line]
        if ctype == "double":
            f.write("  " + ", ".join(repr(float(v)) for v in chunk) + ",\n")
        else:
            f.write("  " + ", ".join(str(int(v)) for v in chunk) + ",\n")
    f.write("};\n\n")


def contnm_tables():
    path = os.path.join(REF, "src", "contnm.f90")
    lines = read_lines(path)
    out = {}

    def grid_table(unit, expect_n):
        blk, l0, l1 = block(lines, r"^block\s+data\s+%s\b" % unit, r"^end\s+block\s+data")
        stmts = join_statements(blk)
        hdr, chunks = None, []
        for name, body in data_statements(stmts):
            if re.match(r"(?i)^v1", name):
                hdr = parse_values(body)
            else:
                chunks.append((name, parse_values(body)))
        return hdr, chunks, (l0, l1)

    for unit, key, n in (("BS296", "SH2O_296", 2003), ("BS260", "SH2O_260", 2003),
                         ("BFH2O", "FH2O", 2003), ("BFCO2", "FCO2", 5003)):
        hdr, chunks, (l0, l1) = grid_table(unit, n)
        names = [c[0].upper() for c in chunks]
        assert names == sorted(names), "DATA chunks out of COMMON order in %s" % unit
        vals = [v for c in chunks for v in c[1]]
        assert len(vals) == n == int(hdr[3]), (unit, len(vals), hdr)
        out[key] = (hdr, vals, "contnm.f90:%d-%d BLOCK DATA %s" % (l0, l1, unit))

    for unit, key in (("BN2T296", "N2RT_296"), ("BN2T220", "N2RT_220")):
        hdr, chunks, (l0, l1) = grid_table(unit, 73)
        d = {c[0].lower(): c[1] for c in chunks}
        ck = [k for k in d if k.startswith("ct")][0]
        sk = [k for k in d if k.startswith("sf")][0]
        assert len(d[ck]) == 73 and len(d[sk]) == 73
        out[key] = (hdr, d[ck], "contnm.f90:%d-%d BLOCK DATA %s (CT)" % (l0, l1, unit))
        out[key + "_SF"] = (hdr, d[sk], "contnm.f90:%d-%d BLOCK DATA %s (sf)" % (l0, l1, unit))

    # XFAC_RHU(-1:61) inside CONTNM, tdep_bandhead inside FRNCO2
    stmts = join_statements(lines[0:1200])
    for name, body in data_statements(stmts):
        if name.upper().startswith("(XFAC_RHU"):
            v = parse_values(body)
            assert len(v) == 63
            out["XFAC_RHU"] = (None, v, "contnm.f90:186-202 XFAC_RHU(-1:61)")
    stmts = join_statements(lines[2950:3020])
    for name, body in data_statements(stmts):
        if name.lower().startswith("(tdep_bandhead"):
            v = parse_values(body)
            assert len(v) == 25
            out["CO2_TDEP_BANDHEAD"] = (None, v, "contnm.f90:2969-2975 tdep_bandhead(1196:1220)")
    assert "XFAC_RHU" in out and "CO2_TDEP_BANDHEAD" in out
    return out


def tips_tables():
    path = os.path.join(REF, "src", "tips_2003.f90")
    lines = read_lines(path)
    stmts = join_statements(lines)
    isonm, tdat = None, None
    for name, body in data_statements(stmts):
        if name.upper().startswith("(ISONM"):
            isonm = [int(v) for v in parse_values(body)]
        if name.lower() == "tdat":
            tdat = parse_values(body)
    assert len(isonm) == 39 and len(tdat) == 119
    # per-molecule QofT tables: walk subroutine by subroutine in file order
    sub_idx = [i for i, l in enumerate(lines) if re.match(r"(?i)^\s*subroutine\s+qt_", l)]
    order = []   # molecule numbers follow the dispatch order 1..38 in TIPS_2003 (lines 64-258)
    tables = {}
    for k, i0 in enumerate(sub_idx):
        i1 = sub_idx[k + 1] if k + 1 < len(sub_idx) else len(lines)
        mol = k + 1
        st = join_statements(lines[i0:i1])
        q = {}
        for name, body in data_statements(st):
            m = re.match(r"(?i)^\(\s*qoft\(\s*(\d+)\s*,\s*j\s*\)\s*,\s*j\s*=\s*1\s*,\s*119\s*\)$", name)
            if m:
                v = parse_values(body)
                assert len(v) == 119, (mol, m.group(1), len(v))
                q[int(m.group(1))] = v
        assert sorted(q) == list(range(1, len(q) + 1)), (mol, sorted(q))
        assert len(q) >= isonm[mol - 1] or mol == 3, (mol, len(q), isonm[mol - 1])
        tables[mol] = q
        order.append(mol)
    assert len(order) == 38
    return isonm, tdat, tables


def smass_table():
    path = os.path.join(REF, "src", "isotope.incl")
    stmts = join_statements(read_lines(path))
    smass = [[0.0] * 9 for _ in range(39)]
    for name, body in data_statements(stmts):
        m = re.match(r"(?i)^\(\s*smass\(\s*(\d+)\s*,\s*i\s*\)\s*,\s*i\s*=\s*1\s*,\s*(\d+)\s*\)$", name)
        if m:
            mol, n = int(m.group(1)), int(m.group(2))
            v = parse_values(body)
            assert len(v) == n
            smass[mol - 1][:n] = v
    assert smass[0][0] == 18.01 and smass[6][0] == 31.99
    return smass


def main():
    os.makedirs(OUT, exist_ok=True)
    ct = contnm_tables()
    with open(os.path.join(OUT, "mtckd_tables.inc"), "w") as f:
        f.write("/* GENERATED by tools/gen_tables.py from the MT_CKD_3.5 DATA tables of the\n"
                " * reference (numbers only).  Do not edit.  Grids: H2O -20..20000 step 10,\n"
                " * CO2 -4..10000 step 2, N2 -10..350 step 5 (cm-1). */\n\n")
        for key in ("SH2O_296", "SH2O_260", "FH2O", "FCO2", "N2RT_296", "N2RT_296_SF",
                    "N2RT_220", "N2RT_220_SF", "XFAC_RHU", "CO2_TDEP_BANDHEAD"):
            hdr, vals, cite = ct[key]
            if hdr is not None:
                f.write("/* grid: V1=%r V2=%r DV=%r NPT=%d */\n" % (hdr[0], hdr[1], hdr[2], int(hdr[3])))
            emit(f, "double", "MTCKD_" + key, vals, cite)

    isonm, tdat, tables = tips_tables()
    with open(os.path.join(OUT, "tips_tables.inc"), "w") as f:
        f.write("/* GENERATED by tools/gen_tables.py from tips_2003.f90 (numbers only). */\n\n")
        emit(f, "int", "TIPS_ISONM", isonm, "tips_2003.f90:370-378 ISONM(39)", per_line=13)
        emit(f, "double", "TIPS_TDAT", tdat, "tips_2003.f90:319-331 Tdat(119)")
        offs, flat = [], []
        nis = []
        for mol in range(1, 39):
            offs.append(len(flat) // 119)
            nis.append(len(tables[mol]))
            for iso in range(1, len(tables[mol]) + 1):
                flat.extend(tables[mol][iso])
        emit(f, "int", "TIPS_QOFFSET", offs, "row offset of molecule m (1..38) in TIPS_QOFT", per_line=13)
        emit(f, "int", "TIPS_QNISO", nis, "number of tabulated isotopologues per molecule", per_line=13)
        f.write("#define TIPS_QROWS %d\n" % (len(flat) // 119))
        emit(f, "double", "TIPS_QOFT", flat, "tips_2003.f90 QofT(iso,1:119), molecules 1..38 in dispatch order")

    smass = smass_table()
    with open(os.path.join(OUT, "smass_table.inc"), "w") as f:
        f.write("/* GENERATED by tools/gen_tables.py from isotope.incl SMASS(39,9); row-major [mol-1][iso-1]. */\n\n")
        emit(f, "double", "ISO_SMASS", [v for row in smass for v in row], "isotope.incl SMASS", per_line=9)
    print("tables written to", os.path.normpath(OUT))


if __name__ == "__main__":
    sys.exit(main())
