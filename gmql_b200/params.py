"""Operator parameters, named after the reference's ADTs so that call sites read like the Scala.

CoverParam N/ANY/ALL:   GMQL-Core/.../DataStructures/CoverParameters/CoverParam.scala:6-16 (+ the parser's ALL/k, (ALL+m)/k
                        closures, Compiler/.../GmqlParsers.scala:543-562)
CoverFlag:              CoverParameters/CoverFlag (COVER, FLAT, SUMMIT, HISTOGRAM)
AtomicCondition:        JoinParametersRD/AtomicCondition.scala:10-30 (DistLess, DistGreater, MinDistance, Upstream, DownStream)
RegionBuilder:          JoinParametersRD/RegionBuilder.scala:7-12
Agg:                    RegionAggregate/RegionsToRegion.scala:9-20 — identified by function_identifier + index
BinSize:                GMQL-Core/.../core/BinSize.scala:14
"""
from dataclasses import dataclass

AGG_KINDS = {"COUNT": 0, "SUM": 1, "MIN": 2, "MAX": 3, "AVG": 4, "MEDIAN": 5, "BAG": 6, "BAGD": 7}


@dataclass(frozen=True)
class Agg:
    function_identifier: str
    index: int = 0            # position of the aggregated value column (RegionsToRegion.index)
    output_name: str = ""

    @property
    def kind(self):
        k = self.function_identifier.upper()
        if k not in AGG_KINDS:
            # DefaultRegionsToRegionFactory.scala:16,29
            raise ValueError("No map function with the given name (%s) found." % self.function_identifier)
        return AGG_KINDS[k]


class CoverFlag:
    COVER, FLAT, SUMMIT, HISTOGRAM = "COVER", "FLAT", "SUMMIT", "HISTOGRAM"
    CODE = {"COVER": 0, "FLAT": 1, "SUMMIT": 2, "HISTOGRAM": 3}


@dataclass(frozen=True)
class CoverParam:
    kind: str          # "N" | "ANY" | "ALL"
    n: int = 0
    add: int = 0       # (ALL + add) / div
    div: int = 1

    def fun(self, x):
        return (x + self.add) // self.div


def N(n):
    return CoverParam("N", n=n)


def ANY():
    return CoverParam("ANY")


def ALL(add=0, div=1):
    return CoverParam("ALL", add=add, div=div)


@dataclass(frozen=True)
class DistLess:
    limit: int


@dataclass(frozen=True)
class DistGreater:
    limit: int


@dataclass(frozen=True)
class MinDistance:
    number: int


@dataclass(frozen=True)
class Upstream:
    pass


@dataclass(frozen=True)
class DownStream:
    pass


ATOM_CODE = {DistLess: 0, DistGreater: 1, MinDistance: 2, Upstream: 3, DownStream: 4}


def atom_tuple(a):
    """(kind code, arg) for the C ABI and (name, arg) for the oracle."""
    if isinstance(a, DistLess):
        return 0, a.limit, "DISTLESS"
    if isinstance(a, DistGreater):
        return 1, a.limit, "DISTGREATER"
    if isinstance(a, MinDistance):
        return 2, a.number, "MD"
    if isinstance(a, Upstream):
        return 3, 0, "UP"
    if isinstance(a, DownStream):
        return 4, 0, "DOWN"
    raise TypeError(a)


class RegionBuilder:
    LEFT, LEFT_DISTINCT, RIGHT, RIGHT_DISTINCT, CONTIG, INTERSECTION, BOTH = (
        "LEFT", "LEFT_DISTINCT", "RIGHT", "RIGHT_DISTINCT", "CONTIG", "INTERSECTION", "BOTH")
    CODE = {"LEFT": 0, "LEFT_DISTINCT": 1, "RIGHT": 2, "RIGHT_DISTINCT": 3, "CONTIG": 4, "INTERSECTION": 5, "BOTH": 6}


@dataclass(frozen=True)
class BinSize:
    Map: int = 50000
    Join: int = 5000
    Cover: int = 2000
