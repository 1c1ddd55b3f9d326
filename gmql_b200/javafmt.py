"""Host-side text forms the reference produces with JVM library calls (strings never cross the C ABI).

java_double_to_string: java.lang.Double.toString — used by the BAG / BAGD aggregates
(GMQL-Server/.../DefaultRegionsToRegionFactory.scala:139,160: `case GDouble(v) => v.toString`).
bag_fold: the pairwise fold of BAG / BAGD as the operators apply it (fun on a two-element list, :133-146 / :154-167).
bag_text: BAG / BAGD applied once to a whole list (GROUP, GroupRD.scala:48-55).
gvalue_text: GValue.toString (GDouble = DecimalFormat("#.########"), RoundingMode.FLOOR; core/DataTypes.scala:139-150,185) —
the text GROUP's grouping fields are compared by (GroupRD.scala:40)."""
import math
from decimal import Context, Decimal, ROUND_FLOOR

_WIDE = Context(prec=400)


def gvalue_text(v) -> str:
    if v is None:
        return "null"
    if not isinstance(v, float):
        return str(v)
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "\u221e" if v > 0 else "-\u221e"
    s = format(Decimal(repr(v)).quantize(Decimal("0.00000001"), rounding=ROUND_FLOOR, context=_WIDE), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s if s not in ("", "-") else "0"


def java_double_to_string(v: float) -> str:
    """Double.toString: decimal notation for 1e-3 <= |v| < 1e7 (at least one digit after the point), otherwise computerised
    scientific notation d.dddE[-]n; the digits are the shortest that identify the double (Python's repr digits)."""
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    sign = "-" if v < 0 else ""
    r = repr(abs(v))
    if "e" in r:
        mant, exp = r.split("e")
        e10 = int(exp)
    else:
        mant, e10 = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    point = len(ip) + e10 - (len(ip + fp) - len((ip + fp).lstrip("0")))     # decimal point position relative to `digits`
    digits = digits.rstrip("0") or "0"
    a = abs(v)
    if 1e-3 <= a < 1e7:
        if point <= 0:
            s = "0." + "0" * (-point) + digits
        elif point >= len(digits):
            s = digits + "0" * (point - len(digits)) + ".0"
        else:
            s = digits[:point] + "." + digits[point:]
        return sign + s
    frac = digits[1:] or "0"
    return "%s%s.%sE%d" % (sign, digits[0], frac, point - 1)


def _cmp_key(v):
    """GValue ordering (core/DataTypes.scala:77-103): GNull < GString < GDouble; strings by compareTo, doubles numerically."""
    if v is None:
        return (0, "", 0.0)
    if isinstance(v, str):
        return (1, v, 0.0)
    return (2, "", v)


def _bag_fun(items, distinct):
    if not items:
        return "."
    if distinct:
        seen = []
        for it in items:
            if it not in seen:
                seen.append(it)
        items = seen
    out = []
    for g in sorted(items, key=_cmp_key):
        out.append("." if g is None else java_double_to_string(g) if isinstance(g, float) else g)
    return ",".join(out)


def bag_fold(values, distinct=False):
    """The operators fold the matched values pairwise (acc = fun([acc, next])), so an output with one contributor keeps
    the raw value and longer lists become comma-joined strings whose layout depends on the fold order; here: the order of
    the experiment rows (the reference's own order is whatever Spark's shuffle produced)."""
    if not values:
        return None
    acc = values[0]
    for v in values[1:]:
        acc = _bag_fun([acc, v], distinct)
    return acc


def bag_text(values, distinct=False):
    """BAG / BAGD on a whole list at once (GROUP applies `fun` to the group's full value list, GroupRD.scala:50)."""
    return _bag_fun(list(values), distinct)
