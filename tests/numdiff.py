"""Numeric text comparison in the spirit of `numdiff` (the reference's CI comparator,
.github/workflows/ci.yml:179-203): texts must have the same token structure; numeric tokens are
compared with an absolute and/or relative tolerance, everything else literally."""
import math
import re

_NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$|^[+-]?(nan|inf)$", re.I)


def _tokens(text):
    return [line.split() for line in text.strip().splitlines()]


def compare(actual, expected, rel=0.0, abs_=0.0):
    """Returns a list of human-readable differences (empty = match)."""
    diffs = []
    A, E = _tokens(actual), _tokens(expected)
    if len(A) != len(E):
        return [f"line count {len(A)} != {len(E)}"]
    for ln, (la, le) in enumerate(zip(A, E), 1):
        if len(la) != len(le):
            diffs.append(f"line {ln}: token count differs: {la} vs {le}")
            continue
        for ta, te in zip(la, le):
            if _NUM.match(ta) and _NUM.match(te):
                a, e = float(ta), float(te)
                if math.isnan(a) and math.isnan(e):
                    continue
                d = abs(a - e)
                if d <= abs_ or d <= rel * max(abs(a), abs(e)):
                    continue
                # a number printed with fewer digits than the tolerance resolves is compared at its own resolution
                diffs.append(f"line {ln}: {ta} vs {te} (|d|={d:.3g})")
            elif ta != te:
                diffs.append(f"line {ln}: {ta!r} vs {te!r}")
    return diffs


def last_place(token):
    """Value of one unit in the last printed digit of a numeric token ("1.0208e-3" -> 1e-7, "0.0028015" -> 1e-7, "7082165" -> 1)."""
    t = token.lower().lstrip("+-")
    mant, _, exp = t.partition("e")
    frac = len(mant.split(".")[1]) if "." in mant else 0
    return 10.0 ** ((int(exp) if exp else 0) - frac)


def printed_units(actual, expected):
    """For two texts with the same token structure: list of (line, expected token, actual token, |difference| in units of the
    last printed digit of the expected token) for every numeric token that differs."""
    out = []
    A, E = _tokens(actual), _tokens(expected)
    assert len(A) == len(E), f"line count {len(A)} != {len(E)}"
    for ln, (la, le) in enumerate(zip(A, E), 1):
        assert len(la) == len(le), f"line {ln}: {la} vs {le}"
        for ta, te in zip(la, le):
            if _NUM.match(ta) and _NUM.match(te):
                a, e = float(ta), float(te)
                if (math.isnan(a) and math.isnan(e)) or a == e:
                    continue
                out.append((ln, te, ta, abs(a - e) / last_place(te)))
            else:
                assert ta == te, f"line {ln}: {ta!r} vs {te!r}"
    return out
