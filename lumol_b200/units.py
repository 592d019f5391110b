"""Unit conversions to lumol's internal units (Angstrom, fs, u, K, mol; energy u*A^2/fs^2).

Host-side convenience mirroring ``lumol_core::units`` (lumol-core/src/units.rs:27-86 for the
factors, :288-406 for ``from`` / ``from_str`` / ``to``) so that tests and inputs read like the
reference's (``units::from(300.0, "kJ/mol/A^2")``).
"""

import math
import re

from .consts import AVOGADRO_NUMBER, BOHR_RADIUS

# Atomic mass unit in kg (units.rs:27)
U_IN_KG = 1.660538782e-27

CONVERSION_FACTORS = {
    # distances
    "A": 1.0, "nm": 10.0, "pm": 1e-2, "fm": 1e-5, "m": 1e10, "bohr": BOHR_RADIUS,
    # time
    "fs": 1.0, "ps": 1e3, "ns": 1e6,
    # mass
    "u": 1.0, "Da": 1.0, "kDa": 1.0, "g": 1e-3 / U_IN_KG, "kg": 1.0 / U_IN_KG,
    # temperature, quantity of matter, angles
    "K": 1.0, "mol": AVOGADRO_NUMBER, "rad": 1.0, "deg": math.pi / 180.0,
    # energy
    "J": 1e-10 / U_IN_KG, "kJ": 1e-7 / U_IN_KG, "kcal": 4.184 * 1e-7 / U_IN_KG,
    "eV": 1.60217653e-19 * 1e-10 / U_IN_KG, "H": 4.35974417e-18 * 1e-10 / U_IN_KG,
    "Ry": 4.35974417e-18 / 2.0 * 1e-10 / U_IN_KG,
    # force
    "N": 1e-20 / U_IN_KG,
    # pressure
    "Pa": 1e-40 / U_IN_KG, "kPa": 1e-37 / U_IN_KG, "MPa": 1e-34 / U_IN_KG, "bar": 1e-35 / U_IN_KG,
    "atm": 101325.0 * 1e-40 / U_IN_KG,
}


class ParseError(ValueError):
    """Error while parsing a unit string (units.rs:88-115)."""


_TOKEN = re.compile(r"\s*([()*/^.]|[^\s()*/^.]+)")


def _tokenize(unit):
    pos, tokens = 0, []
    unit = unit.strip()
    while pos < len(unit):
        match = _TOKEN.match(unit, pos)
        if match is None:
            raise ParseError(f"malformed unit expression: {unit!r}")
        tokens.append(match.group(1))
        pos = match.end()
    return tokens


def _powi(value, power):
    # f64::powi by square-and-multiply, like the reference's UnitExpr::Pow (units.rs:285)
    negative = power < 0
    power = abs(power)
    result = 1.0
    while True:
        if power & 1:
            result *= value
        power //= 2
        if power == 0:
            break
        value *= value
    return 1.0 / result if negative else result


class _Parser:
    def __init__(self, tokens):
        self.tokens = tokens
        self.pos = 0

    def peek(self):
        return self.tokens[self.pos] if self.pos < len(self.tokens) else None

    def next(self):
        token = self.peek()
        self.pos += 1
        return token

    def expr(self):
        value = self.term()
        while self.peek() in ("*", "/", "."):
            op = self.next()
            rhs = self.term()
            value = value / rhs if op == "/" else value * rhs
        return value

    def term(self):
        value = self.factor()
        while self.peek() == "^":
            self.next()
            power = self.next()
            if power is None:
                raise ParseError("Missing value after '^'")
            try:
                value = _powi(value, int(power))
            except ValueError as error:
                raise ParseError(f"Invalid value after ^: {power}") from error
        return value

    def factor(self):
        token = self.next()
        if token is None:
            raise ParseError("missing a value")
        if token == "(":
            value = self.expr()
            if self.next() != ")":
                raise ParseError("Parentheses are not equilibrated.")
            return value
        if token in ")*/^.":
            raise ParseError(f"unexpected '{token}' in unit")
        try:
            return CONVERSION_FACTORS[token]
        except KeyError:
            raise ParseError(f"unit '{token}' not found") from None


def _factor(unit):
    parser = _Parser(_tokenize(unit))
    value = parser.expr()
    if parser.peek() is not None:
        raise ParseError("remaining values after the end of the unit: " + " ".join(parser.tokens[parser.pos:]))
    return value


def from_(value, unit):
    """Convert ``value`` from ``unit`` to internal units (units.rs:372-375)."""
    return _factor(unit) * value


def from_str(value):
    """Parse ``"<number> <unit>"`` (units.rs:384-394)."""
    parts = value.split()
    if not parts:
        raise ParseError("empty value")
    unit = " ".join(parts[1:])
    factor = _factor(unit) if unit else 1.0
    try:
        number = float(parts[0])
    except ValueError as error:
        raise ParseError(f"invalid number {parts[0]!r}") from error
    return factor * number


def to(value, unit):
    """Convert ``value`` from internal units to ``unit`` (units.rs:403-406)."""
    return value / _factor(unit)
