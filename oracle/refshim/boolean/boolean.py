"""Expression tree + recursive-descent parser (see package docstring)."""
import re

_TOKEN = re.compile(r"\s*(?:(?P<op>[&|~!()*+])|(?P<sym>[A-Za-z0-9_][\w.:\-/\[\]=]*))")


class Expression:
    def get_symbols(self):
        out = []
        self._collect(out)
        return out

    def _collect(self, out):
        pass

    def subs(self, mapping):
        return self

    def simplify(self):
        return self

    def _value(self):
        """True / False when the expression is closed, else None."""
        return None

    def __eq__(self, other):
        if not isinstance(other, Expression):
            return NotImplemented
        a, b = self._value(), other._value()
        if a is not None and b is not None:
            return a == b
        return repr(self) == repr(other)

    def __hash__(self):
        return hash(repr(self))


class _Const(Expression):
    def __init__(self, v):
        self.v = bool(v)

    def _value(self):
        return self.v

    def __repr__(self):
        return "TRUE" if self.v else "FALSE"


TRUE = _Const(True)
FALSE = _Const(False)


class Symbol(Expression):
    def __init__(self, obj):
        self.obj = obj

    def _collect(self, out):
        if self not in out:
            out.append(self)

    def subs(self, mapping):
        for k, v in mapping.items():
            if isinstance(k, Symbol) and k.obj == self.obj:
                return v
        return self

    def __str__(self):
        return str(self.obj)

    def __repr__(self):
        return f"Symbol({self.obj!r})"

    def __eq__(self, other):
        return isinstance(other, Symbol) and other.obj == self.obj

    def __hash__(self):
        return hash(("sym", self.obj))


class _Nary(Expression):
    def __init__(self, *args):
        self.args = args

    def _collect(self, out):
        for a in self.args:
            a._collect(out)

    def subs(self, mapping):
        return type(self)(*[a.subs(mapping) for a in self.args])

    def simplify(self):
        v = self._value()
        if v is None:
            return type(self)(*[a.simplify() for a in self.args])
        return TRUE if v else FALSE

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(map(repr, self.args))})"


class AND(_Nary):
    def _value(self):
        vals = [a._value() for a in self.args]
        if any(v is False for v in vals):
            return False
        if all(v is True for v in vals):
            return True
        return None


class OR(_Nary):
    def _value(self):
        vals = [a._value() for a in self.args]
        if any(v is True for v in vals):
            return True
        if all(v is False for v in vals):
            return False
        return None


class NOT(_Nary):
    def _value(self):
        v = self.args[0]._value()
        return None if v is None else (not v)


class BooleanAlgebra:
    TRUE = TRUE
    FALSE = FALSE

    def parse(self, expr, simplify=False):
        if isinstance(expr, Expression):
            return expr
        toks = []
        pos = 0
        expr = expr.strip()
        while pos < len(expr):
            m = _TOKEN.match(expr, pos)
            if not m:
                raise ValueError(f"cannot tokenize {expr!r} at {pos}")
            pos = m.end()
            if m.group("op"):
                toks.append(("op", m.group("op")))
            else:
                s = m.group("sym")
                low = s.lower()
                if low in ("and",):
                    toks.append(("op", "&"))
                elif low in ("or",):
                    toks.append(("op", "|"))
                elif low in ("not",):
                    toks.append(("op", "~"))
                elif low in ("true", "1"):
                    toks.append(("const", True))
                elif low in ("false", "0"):
                    toks.append(("const", False))
                else:
                    toks.append(("sym", s))
        self._toks, self._i = toks, 0
        out = self._or()
        if self._i != len(toks):
            raise ValueError(f"trailing tokens in {expr!r}")
        return out

    def _peek(self):
        return self._toks[self._i] if self._i < len(self._toks) else (None, None)

    def _or(self):
        args = [self._and()]
        while self._peek() in (("op", "|"), ("op", "+")):
            self._i += 1
            args.append(self._and())
        return args[0] if len(args) == 1 else OR(*args)

    def _and(self):
        args = [self._not()]
        while self._peek() in (("op", "&"), ("op", "*")):
            self._i += 1
            args.append(self._not())
        return args[0] if len(args) == 1 else AND(*args)

    def _not(self):
        if self._peek() in (("op", "~"), ("op", "!")):
            self._i += 1
            return NOT(self._not())
        kind, val = self._peek()
        self._i += 1
        if kind == "op" and val == "(":
            e = self._or()
            assert self._peek() == ("op", ")"), "missing )"
            self._i += 1
            return e
        if kind == "const":
            return TRUE if val else FALSE
        if kind == "sym":
            return Symbol(val)
        raise ValueError("unexpected token")
