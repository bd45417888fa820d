"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): executes subroutines of the reference's Fortran sources, read
where they lie under /root/reference (nothing is copied into this repository), through a small source-to-source
translator — Fortran statement by Fortran statement into Python over NumPy arrays with Fortran indexing.

Why: no Fortran compiler, MPI or FFTW exists in this image, so the reference cannot be built.  What CAN be done is to
run its own source text for everything that is plain Fortran: the pointwise loops (calc_flux, calc_rhs, rkt, dealias,
update_uu_prim_from_uu, vardt, update_ksquare, calc_current_density_real ...), the loops around the FFTW calls
(fftw.f90) and the single-rank branches of the transposes (parallel.f90).  The only pieces supplied from outside are
the three FFTW executions (a 1-D DFT each, by definition: numpy.fft on one line) and mpi_allreduce on one rank (a
copy).  tests/golden/make_ref_exec_fixtures.py uses this to produce golden vectors that pin oracle/laps_oracle.py —
and through it the CUDA kernels — to the reference's source instead of to a reading of it.

Supported subset (what those subroutines use): free-form source, `&` continuations, `!` comments, `;` separators,
do / enddo, if / else if / else / endif, one-line if, cycle, exit, call, assignments, array sections with inclusive
bounds, whole-array expressions, parameter declarations, .and./.or./.not./relational operators, cmplx / real / sqrt /
abs / max / min / conjg / aimag / exp / cos / sin / tanh / mod / modulo / floor / int / size.  Everything is FP64 /
complex128, as the reference's -r8 build (makefile: -fdefault-real-8).
"""
from __future__ import annotations

import re

import numpy as np


def _is_int(v):
    return isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_))


class FInt(int):
    """A Fortran INTEGER: closed under + - * and, unlike Python's int, under `/` (integer division truncating toward
    zero, e.g. `normal_size = ngrid / nproc` in decompose_1d, parallel.f90:335).  Mixed with a real it behaves as a real."""

    def _wrap(fn):
        def op(self, o):
            return FInt(fn(int(self), int(o))) if _is_int(o) else fn(int(self), o)
        return op

    def _rwrap(fn):
        def op(self, o):
            return FInt(fn(int(o), int(self))) if _is_int(o) else fn(o, int(self))
        return op

    def _tdiv(a, b):
        if _is_int(a) and _is_int(b):
            q = abs(a) // abs(b)
            return q if (a >= 0) == (b >= 0) else -q
        return a / b

    __add__ = _wrap(lambda a, b: a + b)
    __radd__ = _rwrap(lambda a, b: a + b)
    __sub__ = _wrap(lambda a, b: a - b)
    __rsub__ = _rwrap(lambda a, b: a - b)
    __mul__ = _wrap(lambda a, b: a * b)
    __rmul__ = _rwrap(lambda a, b: a * b)
    __truediv__ = _wrap(_tdiv)
    __rtruediv__ = _rwrap(_tdiv)
    __pow__ = _wrap(lambda a, b: a ** b)
    __mod__ = _wrap(lambda a, b: a % b)

    def __neg__(self):
        return FInt(-int(self))

    def __pos__(self):
        return self


class FArray:
    """A NumPy array indexed the Fortran way: lower bound 1, inclusive upper bounds, first index fastest.
    ``a`` is stored with axes in Fortran order (axis 0 = first Fortran index); wrap a C-ordered [v, z, y, x] array
    with FArray(arr.T) to share memory with it."""

    def __init__(self, a: np.ndarray):
        self.a = a

    def _key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        out = []
        for k in key:
            if isinstance(k, slice):
                lo = None if k.start is None else int(k.start) - 1
                hi = None if k.stop is None else int(k.stop)
                assert k.step is None
                out.append(slice(lo, hi))
            else:
                i = int(k)
                assert i >= 1, "Fortran index below the lower bound"
                out.append(i - 1)
        assert len(out) == self.a.ndim, (len(out), self.a.ndim)
        return tuple(out)

    def __getitem__(self, key):
        v = self.a[self._key(key)]
        return FInt(v) if isinstance(v, np.integer) else v

    def __setitem__(self, key, value):
        self.a[self._key(key)] = value.a if isinstance(value, FArray) else value

    def __iter__(self):
        return iter(self.a)

    def __len__(self):
        return len(self.a)

    # whole-array arithmetic (e.g. `rho_u2 = rho_u2_sum / real(size_grid)`)
    def _v(self, o):
        return o.a if isinstance(o, FArray) else o

    def __add__(self, o): return self.a + self._v(o)
    def __radd__(self, o): return self._v(o) + self.a
    def __sub__(self, o): return self.a - self._v(o)
    def __rsub__(self, o): return self._v(o) - self.a
    def __mul__(self, o): return self.a * self._v(o)
    def __rmul__(self, o): return self._v(o) * self.a
    def __truediv__(self, o): return self.a / self._v(o)
    def __pow__(self, o): return self.a ** o


def _cmplx(a, b=0.0, kind=None):
    return a + 1j * b if np.ndim(a) or np.ndim(b) else complex(a, b)


def _real(x, kind=None):
    if isinstance(x, FArray):
        x = x.a
    if np.ndim(x):
        return np.real(x).astype(np.float64)
    return float(x.real) if isinstance(x, complex) or np.iscomplexobj(x) else float(x)


def _size(x, dim=None):
    a = x.a if isinstance(x, FArray) else np.asarray(x)
    return a.size if dim is None else a.shape[dim - 1]


def _mod(a, p):
    return int(np.fmod(a, p)) if isinstance(a, (int, np.integer)) and isinstance(p, (int, np.integer)) else float(np.fmod(a, p))


INTRINSICS = {
    "cmplx": "_cmplx", "real": "_real", "sqrt": "_np.sqrt", "abs": "abs", "max": "max", "min": "min", "conjg": "_np.conj",
    "aimag": "_np.imag", "exp": "_np.exp", "cos": "_np.cos", "sin": "_np.sin", "tanh": "_np.tanh", "mod": "_mod",
    "modulo": "_modulo", "floor": "_floor", "int": "_int", "size": "_size", "dble": "float", "isnan": "_np.isnan",
    "kind": "_kind",
}

# Library calls whose results come back through scalar (or array-element) arguments: argument positions that are
# outputs.  `call f(a, b, out, ierr)` becomes `out = f(a, b)` (several outputs: a tuple).
OUT_ARGS = {"mpi_comm_rank": (1,), "mpi_comm_size": (1,), "mpi_cart_create": (5,), "mpi_cart_rank": (2,),
            "mpi_comm_split": (3,), "mpi_type_create_subarray": (6,)}

_DECL = re.compile(r"^(integer|real|complex|logical|character|type\s*\(|double\s+precision|use\b|implicit\b|include\b|save\b|external\b)", re.I)
_SKIP = re.compile(r"^(write|print|open|close|inquire|read|allocate|deallocate|format|return\b|contains\b)", re.I)
_IDENT = re.compile(r"[A-Za-z_][A-Za-z_0-9]*")


def _strip_comment(line: str) -> str:
    out, quote = [], None
    for ch in line:
        if quote:
            out.append(ch)
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def statements(text: str):
    """Logical Fortran statements of a source text (comments stripped, continuations joined, `;` split, lower-cased)."""
    stmts, cur = [], ""
    for raw in text.splitlines():
        line = _strip_comment(raw).strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].strip()
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        cur += line
        for part in cur.split(";"):
            part = part.strip()
            if part:
                stmts.append(part.lower())
        cur = ""
    return stmts


def subroutine_statements(path: str, name: str):
    """(dummy argument names, body statements) of ``subroutine name`` in the Fortran source file ``path``."""
    stmts = statements(open(path).read())
    head = re.compile(r"^subroutine\s+%s\b\s*(\((.*)\))?$" % re.escape(name.lower()))
    for i, s in enumerate(stmts):
        m = head.match(s)
        if m:
            args = [a.strip() for a in (m.group(2) or "").split(",") if a.strip()]
            body = []
            for t in stmts[i + 1:]:
                if re.match(r"^end\s*subroutine\b", t):
                    return args, body
                body.append(t)
    raise KeyError(f"subroutine {name} not found in {path}")


def _split_top(s: str, sep: str = ","):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def _match_paren(s: str, i: int) -> int:
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses in: " + s)


class Translator:
    def __init__(self, arrays):
        self.arrays = set(a.lower() for a in arrays)

    def expr(self, s: str) -> str:
        """A whole Fortran expression (operators and literals are rewritten once, then names are resolved recursively)."""
        s = s.strip()
        s = re.sub(r"(\d+\.?\d*|\.\d+)d([+-]?\d+)", r"\1e\2", s)              # 1.0d0 -> 1.0e0
        s = re.sub(r"(?<![\w.])(?<![eE][+-])(\d+)(?![\w.])", r"FInt(\1)", s)       # integer literals are Fortran INTEGERs
        s = re.sub(r"\*\*\s*FInt\((\d+)\)", r"**\1", s)                           # ... except exponents (x**2 stays x*x in NumPy)
        for f, p in ((".and.", " and "), (".or.", " or "), (".not.", " not "), (".true.", " True "), (".false.", " False "),
                     (".eqv.", "=="), (".neqv.", "!="), (".eq.", "=="), (".ne.", "!="), (".gt.", ">"), (".ge.", ">="), (".lt.", "<"), (".le.", "<="), ("/=", "!=")):
            s = s.replace(f, p)
        return self._names(s)

    def _names(self, s: str) -> str:
        out, i = "", 0
        while i < len(s):
            m = _IDENT.match(s, i)
            if m and (i == 0 or not (s[i - 1].isalnum() or s[i - 1] in "_.")):
                name, j = m.group(0), m.end()
                k = j
                while k < len(s) and s[k] == " ":
                    k += 1
                if k < len(s) and s[k] == "(":
                    close = _match_paren(s, k)
                    args = [self._arg(a) for a in _split_top(s[k + 1:close])]
                    if name in self.arrays:
                        out += f"{name}[{', '.join(args)}]"
                    else:
                        out += f"{INTRINSICS.get(name, name)}({', '.join(args)})"
                    i = close + 1
                    continue
                out += name if name not in ("and", "or", "not", "True", "False") else name
                i = j
                continue
            out += s[i]
            i += 1
        return out

    def _arg(self, a: str) -> str:
        parts = _split_top(a, ":")
        if len(parts) == 1:
            return self._names(a.strip())
        return ":".join(self._names(p.strip()) if p.strip() else "" for p in parts)

    def subroutine(self, name: str, args, body, py_name=None) -> str:
        """Python source of one subroutine.  Names that are assigned but neither declared in the subroutine nor dummy
        arguments are module variables (`global`)."""
        local = set(args)
        local_arrays = set()
        lines, indent = [], 1
        assigned = set()
        saved_arrays = set(self.arrays)
        local_alloc = set()
        for st in body:          # dummies and locals declared with a shape are arrays inside this subroutine
            if _DECL.match(st) and "::" in st:
                attrs, names = st.split("::", 1)
                for item in _split_top(names):
                    nm = _IDENT.match(item.strip()).group(0)
                    if "dimension" in attrs or ("(" in item.split("=")[0]):
                        self.arrays.add(nm)
                        if nm not in args:
                            local_alloc.add(nm)

        def emit(t):
            lines.append("    " * indent + t)

        for st in body:
            if _DECL.match(st):
                if "::" in st:
                    attrs, names = st.split("::", 1)
                    for item in _split_top(names):
                        item = item.strip()
                        nm = _IDENT.match(item).group(0)
                        local.add(nm)
                        if "parameter" in attrs and "=" in item:
                            emit(f"{nm} = {self.expr(item.split('=', 1)[1])}")
                        elif "=" in item:                          # initialised local (implicit save): plain local here
                            emit(f"{nm} = {self.expr(item.split('=', 1)[1])}")
                        elif "(" in item or "dimension" in attrs:
                            local_arrays.add(nm)
                            # explicit shape: `real, dimension(3) :: d` / `integer :: a(3)`; deferred shape waits for allocate
                            shape = item[item.index("(") + 1:_match_paren(item, item.index("("))] if "(" in item else None
                            if shape is None:
                                m2 = re.search(r"dimension\s*\(", attrs)
                                if m2:
                                    k2 = attrs.index("(", m2.start())
                                    shape = attrs[k2 + 1:_match_paren(attrs, k2)]
                            if shape is not None and ":" not in shape and nm not in args:
                                dt = "complex" if attrs.strip().startswith("complex") else ("int" if attrs.strip().startswith("integer") else "float")
                                emit(f"{nm} = _alloc(({self.expr(shape)},), '{dt}')")
                else:                                              # old-style: `real dt` / `integer irk`
                    for item in _split_top(re.sub(r"^\w+(\s*\*\s*\d+)?\s+", "", st)):
                        m = _IDENT.match(item.strip())
                        if m:
                            local.add(m.group(0))
                continue
            if st.startswith("allocate"):                         # local allocatables only; module arrays are set up by the caller
                k = st.index("(")
                for item in _split_top(st[k + 1:_match_paren(st, k)]):
                    item = item.strip()
                    nm = _IDENT.match(item).group(0)
                    if nm in local_alloc and "(" in item:
                        shape = item[item.index("(") + 1:_match_paren(item, item.index("("))]
                        emit(f"{nm} = _alloc(({self.expr(shape)},), 'float')")
                continue
            if _SKIP.match(st):
                if st.startswith("return"):
                    emit("return")
                continue
            m = re.match(r"^do\s+(\w+)\s*=\s*(.+)$", st)
            if m:
                rng = [self.expr(p) for p in _split_top(m.group(2))]
                step = rng[2] if len(rng) > 2 else "1"
                emit(f"for {m.group(1)} in _frange({rng[0]}, {rng[1]}, {step}):")
                assigned.add(m.group(1))
                indent += 1
                continue
            if re.match(r"^end\s*do$", st) or re.match(r"^end\s*if$", st):
                indent -= 1
                continue
            m = re.match(r"^(else\s*)?if\s*\(", st)
            if m:
                k = st.index("(", m.start())
                close = _match_paren(st, k)
                cond, rest = self.expr(st[k + 1:close]), st[close + 1:].strip()
                if rest == "then":
                    if m.group(1):
                        indent -= 1
                        emit(f"elif {cond}:")
                    else:
                        emit(f"if {cond}:")
                    indent += 1
                    emit("pass")
                else:                                              # one-line if
                    emit(f"if {cond}:")
                    indent += 1
                    self._simple(rest, emit, assigned)
                    indent -= 1
                continue
            if st == "else":
                indent -= 1
                emit("else:")
                indent += 1
                emit("pass")
                continue
            self._simple(st, emit, assigned)
        self.arrays = saved_arrays
        glob = sorted(n for n in assigned if n not in local)
        head = [f"def {py_name or name}({', '.join(args)}):"]
        if glob:
            head.append("    global " + ", ".join(glob))
        head.append("    pass")
        return "\n".join(head + lines) + "\n"

    def _simple(self, st, emit, assigned):
        if st == "cycle":
            emit("continue")
        elif st == "exit":
            emit("break")
        elif st.startswith("call mpi_allreduce"):                  # one rank: the "sum over ranks" is the local value
            a, b = [x.strip() for x in _split_top(st[st.index("(") + 1:_match_paren(st, st.index("("))])[:2]]
            if b in self.arrays:
                emit(f"{b}.a[...] = {a}.a")
            else:
                assigned.add(b)
                emit(f"{b} = {a}")
        elif st.startswith("call ") and _IDENT.match(st[5:].strip()).group(0) in OUT_ARGS:
            rest = st[5:].strip()
            nm = _IDENT.match(rest).group(0)
            k = rest.index("(")
            args = [a.strip() for a in _split_top(rest[k + 1:_match_paren(rest, k)])]
            outs = [args[i] for i in OUT_ARGS[nm]]
            ins = [self.expr(a) for i, a in enumerate(args[:-1]) if i not in OUT_ARGS[nm]]      # the last argument is ierr
            for o in outs:
                if "(" not in o:
                    assigned.add(o)
            emit(f"{', '.join(self.expr(o) for o in outs)} = {nm}({', '.join(ins)})")
        elif st.startswith("call "):
            rest = st[5:].strip()
            emit(self.expr(rest) if "(" in rest else f"{rest}()")
        else:
            # assignment: the first top-level '=' that is not part of ==, <=, >=, /=
            depth = 0
            for i, ch in enumerate(st):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "=" and depth == 0 and st[i + 1:i + 2] != "=" and st[i - 1] not in "=<>/":
                    lhs, rhs = st[:i].strip(), st[i + 1:]
                    nm = _IDENT.match(lhs).group(0)
                    if "(" not in lhs:
                        if nm in self.arrays:                      # whole-array assignment keeps the storage
                            emit(f"{nm}.a[...] = {self.expr(rhs)}")
                        else:
                            assigned.add(nm)
                            emit(f"{nm} = {self.expr(rhs)}")
                    else:
                        emit(f"{self.expr(lhs)} = {self.expr(rhs)}")
                    return
            raise ValueError("cannot translate statement: " + st)


def _alloc(shape, kind):
    dt = {"float": np.float64, "int": np.int64, "complex": np.complex128}[kind]
    return FArray(np.zeros(tuple(int(n) for n in shape)[::-1], dtype=dt).T)


def case_body(body, selector: str, value: str):
    """The statements of `case(value)` of the (outermost) `select case(selector)` in a subroutine body."""
    out, depth, active, inside = [], 0, False, False
    for st in body:
        m = re.match(r"^select\s*case\s*\((.*)\)$", st)
        if m:
            depth += 1
            if depth == 1 and m.group(1).strip() == selector:
                inside = True
                continue
        elif re.match(r"^end\s*select", st):
            depth -= 1
            if depth == 0 and inside:
                return out
        elif inside and depth == 1 and re.match(r"^case\b", st):
            m = re.match(r"^case\s*\((.*)\)$", st)
            active = bool(m) and value in [v.strip() for v in m.group(1).split(",")]
            continue
        if inside and active:
            out.append(st)
    raise KeyError(f"select case({selector}) / case({value}) not found")


def _frange(a, b, step=1):
    return (FInt(i) for i in range(int(a), int(b) + (1 if step > 0 else -1), int(step)))


def base_namespace():
    return {"_np": np, "_cmplx": _cmplx, "_real": _real, "_size": _size, "_mod": _mod, "_frange": _frange,
            "_modulo": lambda a, p: a % p, "_floor": lambda x: FInt(np.floor(x)), "_kind": lambda x: 8, "_int": lambda x: FInt(x),
            "FInt": FInt, "_alloc": _alloc}


def load(namespace: dict, path: str, names, arrays=None):
    """Translate the subroutines ``names`` of the Fortran file ``path`` and define them in ``namespace``.  Array names
    are the FArray objects already in the namespace (plus ``arrays``)."""
    arr = {k for k, v in namespace.items() if isinstance(v, FArray)} | set(arrays or ())
    tr = Translator(arr)
    src = {}
    for n in names:
        args, body = subroutine_statements(path, n)
        code = tr.subroutine(n.lower(), args, body)
        src[n] = code
        exec(compile(code, f"<{path}:{n}>", "exec"), namespace)
    return src
